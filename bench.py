#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: block-sparse contraction useful FP64 GFLOP/s and % of the
measured FP64 tensor-core GEMM peak.

Workload at N=1 (config.workload): BASELINE configs[2], the headline -- U(1) two-site
effective-Hamiltonian apply at D=4096, complex double (four chained Contract calls).
One "step" = one H_eff apply.  `value` = algorithmic flops (the reference's own cost model,
tensor_op_cost.h: 8*m*k*n per complex pair) / device time with all operands resident in HBM.
`e2e` = the same apply through the public chain API with psi in pinned HOST memory: H2D of psi,
four steps, D2H of the result inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--D 4096] [--dtype c128|f64]
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEEDS = {"f64": 20260002, "c128": 20260003}



def _fatal(msg):
    """Verification failure: fatal, except in kernel attribution experiments (exp/, QLB200_BENCH_NO_VERIFY=1) that run
    deliberately wrong kernel variants for their timing only."""
    if os.environ.get("QLB200_BENCH_NO_VERIFY") == "1":
        print("[not verified] " + msg, file=sys.stderr)
        return
    raise SystemExit(msg)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="heff_u1", choices=["heff_u1", "heff_hubbard", "ragged"],
                    help="heff_u1: BASELINE configs[1]/[2] (default, the headline); heff_hubbard: configs[3] (U(1)xU(1) fermionic "
                         "Hubbard H_eff apply, default D=8192 double); ragged: configs[4] (10^4 random blocks, permute + grouped GEMM)")
    ap.add_argument("--D", type=int, default=None, help="bond dimension (default 4096; 8192 for heff_hubbard)")
    ap.add_argument("--dtype", default=None, choices=["c128", "f64"], help="default c128; f64 for heff_hubbard and ragged")
    ap.add_argument("--cpu-sample-D", type=int, default=0, help="reference arm / cpu_baseline: bond dimension of the CPU run (0 = the workload's own D, i.e. the same configuration)")
    ap.add_argument("--cpu-time-cap", type=float, default=150.0, help="reference arm: stop adding timed applies after this many seconds (steps actually run are reported)")
    ap.add_argument("--sweep-D", type=int, default=2048, help="reference arm: bond dimension of the thread sweep")
    ap.add_argument("--no-thread-sweep", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=3, help="GPU arm: timed applies of the cpu_baseline child")
    ap.add_argument("--write-out", default="", help="reference arm: write the result tensor of the apply to this file (the reference's stream format)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--plumbing", default="capi", choices=["capi", "torch"],
                    help="N>1: symmetric buffers / multicast mapping / barrier from qlb200_comm_* (C ABI, default) or from torch symmetric memory + CUDA IPC")
    ap.add_argument("--rebalance", type=int, default=4, help="N>1: feedback iterations of the row partitioner (measure every rank's share, re-weigh, cut again)")
    ap.add_argument("--rebalance-damp", type=float, default=0.5, help="N>1: fraction (exponent) of the feedback correction applied per iteration")
    ap.add_argument("--snap", type=int, default=8, help="N>1: row cuts inside a sector are multiples of this")
    ap.add_argument("--no-sub-records", action="store_true", help="default run only: skip the sub-records for BASELINE configs[1], [3] and [4]")
    ap.add_argument("--ragged-cpu-pairs", type=int, default=1000, help="ragged workload: pairs of the bounded CPU / e2e sample")
    ap.add_argument("--no-fused-mpo", action="store_true", help="skip the measurement of the chain with the two MPO tensors pre-contracted")
    ap.add_argument("--no-cold", action="store_true", help="skip the cold drop-in measurement (4 Contract calls on host tensors incl. match + plan build)")
    ap.add_argument("--pipe-in", default="", help="N=1 e2e: cumulative fractions of the streamed input at which the parts of step 1 are cut (tuning; default ContractionChain.pipe_fractions)")
    ap.add_argument("--pipe-out", default="", help="N=1 e2e: cumulative fractions of the output at which the parts of the last step are cut (tuning)")
    ap.add_argument("--self-check", action="store_true", help="N=1: compare the apply with the same apply through the alternate kernels of the library (reference-free; for A/B runs with --no-cpu-baseline)")
    ap.add_argument("--breakdown", action="store_true", help="print the per-kernel table to stderr")
    ap.add_argument("--plan-flags", type=int, default=1, help="qlb200_plan_create flags (kernel A/B testing; 1 = default)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "multicast", "fused", "allgather"],
                    help="N>1: output exchange fused into the last GEMM's epilogue -- multimem.st through the NVSwitch multicast mapping "
                         "(multicast; auto picks it when the fabric offers it) or unicast peer stores (fused) -- or NCCL all-gather + unpack")
    ap.add_argument("--tensors", default="", help="directory with lenv.qlten, psi.qlten, mpo1.qlten, mpo2.qlten, renv.qlten written by a "
                                                  "TensorToolkit program (operator<<): run the H_eff apply on those tensors instead of synthetic ones")
    ap.add_argument("--qn", default="U1QN", choices=["U1QN", "fU1QN", "U1U1QN", "fU1U1QN", "Z2QN", "fZ2QN"], help="quantum-number type of --tensors files")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel of an apply from the host instead of replaying one CUDA graph")
    ap.add_argument("--shard-of", default="", help="W:r -- time rank r's share of a W-GPU run on one GPU, no collective (tuning aid)")
    args = ap.parse_args()
    if args.D is None:
        args.D = 8192 if args.workload == "heff_hubbard" else 4096
    if args.dtype is None:
        args.dtype = "c128" if args.workload == "heff_u1" else "f64"
    return args


def np_dtype(name):
    return np.complex128 if name == "c128" else np.float64


def workload_name(D, dtype, workload="heff_u1"):
    model = ("U(1)xU(1) fermionic Hubbard (Grassmann tensors, fU1U1QN)" if workload == "heff_hubbard" else "U(1) spin-1/2 Heisenberg")
    return (f"{model} two-site effective-Hamiltonian apply (lenv x psi x W1 x W2 x renv, 4 chained Contract), "
            f"D={D}, {'complex double' if dtype == 'c128' else 'double'}")


def workload_indexes(workload, D):
    from tensortoolkit_b200 import workloads as wl
    return wl.hubbard_indexes(D) if workload == "heff_hubbard" else wl.u1_heisenberg_indexes(D)


def workload_seed(workload, dtype):
    return 20260004 if workload == "heff_hubbard" else SEEDS[dtype]


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and clock-event (throttle) reasons sampled DURING the timed region (B200_PROFILING.md recipe).
    In-process NVML thread, one sample every 20 ms, so that even a 30 ms multi-GPU timed region is covered; falls back
    to an `nvidia-smi -lms 100` child process when NVML cannot be initialised."""
    _REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index, self.proc, self.samples = index, None, []      # samples: (sm_mhz, sm_max_mhz, power_w, reason bitmask)
        self.thread, self.stop_flag, self.nvml = None, None, None

    def _nvml_loop(self):
        n = self.nvml
        try:
            h = n.nvmlDeviceGetHandleByIndex(self.index)
            smax = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
            reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
            while True:
                try:
                    pw = n.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = None
                self.samples.append((float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)), smax, pw, int(reasons_fn(h))))
                if self.stop_flag.wait(0.02):
                    break
        except Exception:
            pass

    def start(self):
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.stop_flag = pynvml, threading.Event()
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = self.thread = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            time.sleep(0.03)
            self.stop_flag.set()
            self.thread.join(timeout=2)
            self.thread = None
            return
        if self.proc is None:
            return
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        num = lambda x: float(x) if x.replace(".", "", 1).isdigit() else None
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) >= 7:
                mask = sum(bit for (bit, _), v in zip(self._REASONS, f[3:7]) if v.lower().startswith("active"))
                self.samples.append((num(f[0]), num(f[1]), num(f[2]), mask))

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(s[0] for s in self.samples if s[0] is not None)
        pw = [s[2] for s in self.samples if s[2] is not None]
        reasons = [name for bit, name in self._REASONS if any(s[3] & bit for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.samples[0][1],
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(self.samples)}


def load_tensors(directory, qn_name, dtype):
    """Real tensors in the reference's file format (tensortoolkit_b200/qlten_io.py); indexes must chain like workloads.HEFF_STEPS."""
    import tensortoolkit_b200 as tk
    from tensortoolkit_b200 import qlten_io
    kind = {k.name: k for k in (tk.U1, tk.fU1, tk.U1U1, tk.fU1U1, tk.Z2, tk.fZ2)}[qn_name]
    return {name: qlten_io.load(os.path.join(directory, name + ".qlten"), kind, np_dtype(dtype)) for name in ("lenv", "psi", "mpo1", "mpo2", "renv")}


def build_tensors(D, dtype, rng, workload="heff_u1"):
    import tensortoolkit_b200 as tk
    from tensortoolkit_b200 import workloads as wl
    ti = wl.heff_tensor_indexes(workload_indexes(workload, D))
    div = (0,) * ti["psi"][0].kind.nvals
    return {name: tk.BlockSparseTensor(idxs, np_dtype(dtype)).random(div, rng) for name, idxs in ti.items()}


# ------------------------------------------------------------------------------------------------
class RefApply:
    """The reference's own CPU path (oracle/_ref/libqlref.so: TensorToolkit + HPTT + OpenBLAS) on one H_eff apply:
    operands built once (the reference's seeded Random, or files written in its stream format), every apply = the four
    chained qlten::Contract calls of workloads.HEFF_STEPS, Contract calls only inside the timed region."""

    def __init__(self, D, dtype, workload="heff_u1", tensors_dir="", qn="U1QN"):
        from oracle import refbridge as ref
        from tensortoolkit_b200 import workloads as wl
        self.ref, self.wl = ref, wl
        ref.lib()
        ti = wl.heff_tensor_indexes(workload_indexes(workload, D))
        if tensors_dir:
            import tensortoolkit_b200 as tk
            from tensortoolkit_b200 import qlten_io
            kind = {k.name: k for k in (tk.U1, tk.fU1, tk.U1U1, tk.fU1U1, tk.Z2, tk.fZ2)}[qn]
            self.r = {}
            for name in ("lenv", "psi", "mpo1", "mpo2", "renv"):
                path = os.path.join(tensors_dir, name + ".qlten")
                idxs = qlten_io.load(path, kind, np_dtype(dtype)).indexes       # our reader: indexes only (to type the handle)
                self.r[name] = ref.RefTensor.new(idxs, np_dtype(dtype)).read_file(path)   # the reference's own reader
        else:
            ref.set_seed(workload_seed(workload, dtype))
            div = (0,) * ti["psi"][0].kind.nvals
            self.r = {name: ref.RefTensor.new(idxs, np_dtype(dtype)).random(div) for name, idxs in ti.items()}
        self.flops = 0.0
        for lhs, rhs, axes, out in wl.HEFF_STEPS:       # builds the intermediates; doubles as the warm-up pass
            self.flops += ref.contract_cost(self.r[lhs], self.r[rhs], axes)["flops"]
            self.r[out] = ref.contract(self.r[lhs], self.r[rhs], axes)

    def apply(self, threads):
        """Seconds of one apply (sum of the four Contract calls)."""
        self.ref.set_threads(threads)
        return sum(self.ref.contract_time(self.r[lhs], self.r[rhs], axes, 1) for lhs, rhs, axes, _ in self.wl.HEFF_STEPS)


def pick_threads():
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun injects OMP_NUM_THREADS=1; the reference arm uses every host core it can (set before the BLAS / OpenMP
    # runtimes are loaded, and again explicitly through the reference's own SetTensorManipulationThreads)
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "GOTO_NUM_THREADS"):
        os.environ.pop(k, None)
    threads = pick_threads()
    D = args.cpu_sample_D or args.D          # default: the SAME configuration as the GPU arm
    t_setup = time.perf_counter()
    ra = RefApply(D, args.dtype, args.workload, args.tensors, args.qn)
    times = []
    # calibration (doubles as warm-up): all host threads vs half of them -- HPTT's OpenMP pool and OpenBLAS's pthread pool
    # oversubscribe on some boxes; the timed run uses whichever is faster, `cores` says which
    all_threads = threads
    if threads >= 4:
        cal = {t: ra.apply(t) for t in (threads, threads // 2)}
        threads = min(cal, key=cal.get)
    for _ in range(max(0, args.warmup - 2)):          # RefApply's constructor already ran one full apply
        ra.apply(threads)
        if time.perf_counter() - t_setup > args.cpu_time_cap / 2:
            break
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        times.append(ra.apply(threads))
        if time.perf_counter() - t0 > args.cpu_time_cap:       # bounded sample: report the steps actually run
            break
    sec = float(np.mean(times))
    val = ra.flops / sec / 1e9
    if args.write_out:
        ra.r["out"].write_file(args.write_out)             # the reference's result, in its own stream format (parity check)
    # thread sweep: best-t vs all-cores, on a smaller bond dimension so that t = 1 stays bounded
    sweep = None
    if not args.no_thread_sweep:
        Ds = min(D, args.sweep_D)
        rs = ra if (Ds == D and not args.tensors) else RefApply(Ds, args.dtype, args.workload)
        sweep = {"D": Ds, "gflops_by_threads": {}}
        t = 1
        cand = []
        while t < all_threads:
            cand.append(t); t *= 2
        cand.append(all_threads)
        budget = time.perf_counter() + args.cpu_time_cap / 2
        for t in reversed(cand):                  # most threads first; stop when the budget is gone
            if time.perf_counter() > budget:
                break
            best = min(rs.apply(t) for _ in range(2))
            sweep["gflops_by_threads"][str(t)] = rs.flops / best / 1e9
        bt = max(sweep["gflops_by_threads"], key=lambda k: sweep["gflops_by_threads"][k])
        sweep["best_threads"], sweep["best_gflops"] = int(bt), sweep["gflops_by_threads"][bt]
    same = D == args.D
    sample = (f"{len(times)} H_eff applies at D={D} ({ra.flops / 1e9:.1f} GFLOP each"
              + ("" if same else f"; the D={args.D} workload is {(args.D / D) ** 3:.0f}x the flops") + "), Contract calls only, "
              f"{threads} of {all_threads} host threads (the faster of all / half, OMP_WAIT_POLICY=passive)")
    line = {
        "impl": "reference", "metric": "block-sparse contraction useful FP64 GFLOP/s", "value": val, "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "c64(f64 pairs)" if args.dtype == "c128" else "f64",
        "data": "files" if args.tensors else "synthetic",
        "config": {"workload": workload_name(args.D, args.dtype, args.workload), "bounded_sample": sample, "same_config": same,
                   "steps_requested": args.steps},
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": threads, "kind": "reference", "sample": sample, "thread_sweep": sweep},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def measure_fp64_peak(torch, dtype):
    """cuBLAS GEMM peak for the arithmetic type, measured here (MEASURED_PEAKS.json has no FP64 entry)."""
    n = 4096 if dtype == "c128" else 8192
    tdt = torch.complex128 if dtype == "c128" else torch.float64
    a = torch.randn(n, n, dtype=tdt, device="cuda")
    b = torch.randn(n, n, dtype=tdt, device="cuda")
    fl = (8.0 if dtype == "c128" else 2.0) * n ** 3
    best = float("inf")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    # sustained: back to back for ~1.5 s
    reps = max(3, int(1.5 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b)
    e1.record(); torch.cuda.synchronize()
    sustained = fl * reps / (e0.elapsed_time(e1) * 1e-3)
    del a, b
    return fl / best / 1e12, sustained / 1e12, f"torch.matmul {'complex128' if dtype == 'c128' else 'float64'} {n}^3 (cuBLAS), best of 6 / back-to-back {reps} calls"


class Env:
    """One process = one GPU: torch for device memory / streams / the NCCL process group, one qlb200 context on its own stream."""

    def __init__(self):
        import torch
        import tensortoolkit_b200 as tk
        self.torch, self.tk = torch, tk
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        torch.cuda.set_device(self.local)
        self.stream = torch.cuda.Stream()
        self.ctx = tk.Context(self.local)
        self.ctx.set_stream(self.stream.cuda_stream)
        self.flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2
        self._peaks = {}

    def fp64_peak(self, dtype):
        if dtype not in self._peaks:
            with self.torch.cuda.stream(self.stream):
                self._peaks[dtype] = measure_fp64_peak(self.torch, dtype)
        return self._peaks[dtype]

    def shared_host_buffer(self, name, nbytes):
        """Host memory that every rank of the node maps (POSIX shared memory): the ranks' downloads of their own result
        slabs land in ONE buffer.  Returns (numpy uint8 view, handle to keep alive)."""
        if self.world == 1:
            a = np.empty(nbytes, np.uint8)
            return a, None
        from multiprocessing import shared_memory
        import torch.distributed as dist
        tag = f"qlb200_{os.environ.get('MASTER_PORT', '0')}_{name}"
        shm = None
        if self.rank == 0:
            try:
                shared_memory.SharedMemory(name=tag).unlink()      # stale segment of a crashed run
            except Exception:
                pass
            shm = shared_memory.SharedMemory(name=tag, create=True, size=max(nbytes, 1))
        dist.barrier()
        if self.rank != 0:
            shm = shared_memory.SharedMemory(name=tag)
            try:                                                    # only the creator unlinks
                from multiprocessing import resource_tracker
                resource_tracker.unregister(shm._name, "shared_memory")
            except Exception:
                pass
        return np.ndarray((nbytes,), np.uint8, buffer=shm.buf), shm


def measure_heff(args, env):
    """One record for an H_eff workload (args.workload / D / dtype) at env.world GPUs: device-timed value, per-kernel
    roofline, end-to-end through host memory, and at one GPU the reference CPU arm + parity against it."""
    torch, tk = env.torch, env.tk
    from tensortoolkit_b200 import workloads as wl
    from tensortoolkit_b200.heff import ContractionChain
    rank, world, local, stream, ctx = env.rank, env.world, env.local, env.stream, env.ctx

    dtype = args.dtype
    es = 16 if dtype == "c128" else 8
    rng = np.random.default_rng(workload_seed(args.workload, dtype))
    tensors = load_tensors(args.tensors, args.qn, dtype) if args.tensors else build_tensors(args.D, dtype, rng, args.workload)
    sharded = None
    if args.shard_of:
        from tensortoolkit_b200.heff import ShardedChain
        ew, er = (int(x) for x in args.shard_of.split(":"))
        with torch.cuda.stream(stream):
            sharded = ShardedChain(ctx, tensors, wl.HEFF_STEPS, "lenv", 2, np_dtype(dtype), ew, er, flags=args.plan_flags, exchange=None)
        chain = sharded.chain
        apply_fn = sharded.apply
    elif world > 1:
        # partition by output sector / row slab of lenv's free bond; all-gather of disjoint slabs per apply
        from tensortoolkit_b200.heff import ShardedChain
        with torch.cuda.stream(stream):
            sharded = ShardedChain(ctx, tensors, wl.HEFF_STEPS, "lenv", 2, np_dtype(dtype), world, rank, flags=args.plan_flags,
                                   exchange=args.exchange, host_input="psi", snap=args.snap, plumbing=args.plumbing)
        # partitioner feedback: time every rank's share, re-weigh the row line by (measured time / modelled flops) per rank,
        # cut again (sharding.reweigh_pieces) -- an autotuning step at set-up, like the planner's split-K simulation
        rebalance_log = []
        from tensortoolkit_b200 import sharding as shd

        def build(pieces):
            with torch.cuda.stream(stream):
                return ShardedChain(ctx, tensors, wl.HEFF_STEPS, "lenv", 2, np_dtype(dtype), world, rank, flags=args.plan_flags,
                                    exchange=args.exchange, host_input="psi", pieces=pieces, snap=args.snap, plumbing=args.plumbing)

        def local_times(sharded):
            """Median local-step time of every rank under the current cuts (host-launched applies; the launch overhead is the
            same on every rank and only damps the correction)."""
            with torch.cuda.stream(stream):
                ts = []
                for i in range(12):
                    env.flush.zero_()
                    torch.distributed.barrier()
                    marks = {}
                    e0 = torch.cuda.Event(enable_timing=True); e0.record(stream)
                    def mark(label, marks=marks):
                        ev = torch.cuda.Event(enable_timing=True); ev.record(stream); marks[label] = ev
                    sharded.apply(mark)
                    torch.cuda.synchronize()
                    if i >= 2:
                        ts.append(e0.elapsed_time(marks["compute"]))
                mine_t = torch.tensor([float(np.median(ts))], device="cuda", dtype=torch.float64)
                allt = [torch.zeros_like(mine_t) for _ in range(world)]
                torch.distributed.all_gather(allt, mine_t)
                return [float(x[0]) for x in allt]

        # every candidate cut is measured and the best kept (sharding.tune_partition); all ranks see the same gathered times
        rebalance_kept = 0
        if args.rebalance:
            sharded, rebalance_kept, log = shd.tune_partition(sharded, build, local_times, rounds=args.rebalance, damp=args.rebalance_damp)
            rebalance_log = [[round(t, 4) for t in times] for times in log]
        chain = sharded.chain
        apply_fn = sharded.apply
    else:
        chain = ContractionChain(ctx, tensors, wl.HEFF_STEPS, np_dtype(dtype), args.plan_flags)
        apply_fn = chain.apply_device
    verified = None
    if world > 1:
        # parity of the sharded path on this very input: every rank rebuilds the full result and checks it
        # against its own unsharded apply (relative Frobenius error, tolerance 1e-12)
        with torch.cuda.stream(stream):
            ref_chain = ContractionChain(ctx, tensors, wl.HEFF_STEPS, np_dtype(dtype), args.plan_flags)
            ref_chain.apply_device()
            want = ref_chain.result("out").data
            ref_chain.close()
            tk._lib.check(tk._lib.lib.qlb200_memcpy_h2d(ctx.h, sharded.in_ptr, tensors["psi"].data.ctypes.data, tensors["psi"].data.nbytes), "h2d")
            sharded.apply()
            torch.distributed.barrier()
            got = np.empty_like(want)
            sharded.download_full(got)
            ctx.sync()
        verified = float(np.linalg.norm(got - want) / np.linalg.norm(want))
        if not verified <= 1e-12:
            _fatal(f"rank {rank}: sharded apply differs from the unsharded one, rel err {verified:.3e}")
        del got
    self_check = None
    if args.self_check and world == 1 and not args.shard_of:
        # reference-free check for kernel A/B runs that skip the CPU baseline (and with it the parity key): the same apply through
        # DIFFERENT kernels of this library -- every block through the permute kernel, the four-product complex GEMM, no
        # narrow-pair kernel -- must give the same tensor.  (A variant that skips or repeats work is wrong in only one of the two.)
        alt_flags = (tk._lib.PLAN_DETERMINISTIC | tk._lib.PLAN_NO_SKINNY | tk._lib.PLAN_PERMUTE_ALL |
                     (tk._lib.PLAN_CPLX_4M if dtype == "c128" else 0))
        with torch.cuda.stream(stream):
            alt = ContractionChain(ctx, tensors, wl.HEFF_STEPS, np_dtype(dtype), alt_flags)
            alt.apply_device()
            want_alt = alt.result("out").data
            alt.close()
            chain.apply_device()
            got_alt = chain.result("out").data
        sc = float(np.linalg.norm(got_alt - want_alt) / np.linalg.norm(want_alt))
        if not sc <= 1e-12:
            _fatal(f"self-check FAILED: this plan's kernels and the alternate kernel path disagree, rel err {sc:.3e}")
        self_check = {"rel_err_vs_alternate_kernels": sc, "alternate": "PLAN_NO_SKINNY | PLAN_PERMUTE_ALL" + (" | PLAN_CPLX_4M" if dtype == "c128" else ""),
                      "tolerance": 1e-12}
        del want_alt, got_alt
    graph = None
    if not args.no_graph:
        # one apply = one CUDA graph launch (kernels, and for N>1 the exchange barrier): no per-kernel host launch cost
        with torch.cuda.stream(stream):
            graph = (sharded if sharded is not None else chain).capture()
        apply_eager, apply_fn = apply_fn, graph.launch
    stats = chain.stats()
    flops_local = chain.flops()
    flush = env.flush

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(max(3, args.warmup)):
            apply_fn()
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        # ---- timed region: K applies, per-step events, L2 flushed between steps ----
        evs = []
        for _ in range(args.steps):
            flush.zero_()
            if sharded is not None and world > 1:
                sharded.barrier()     # the ranks leave their L2 flushes at different times: start every timed apply together
                                      # (stream-ordered, before the first event: not part of the timed region)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            launches = apply_fn()
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        dev_ms = [a.elapsed_time(b) for a, b in evs]
        # ---- per-kernel breakdown (same stream, CUDA events around each launch group) ----
        kern = []
        for si, ((lhs, rhs, axes, out), plan) in enumerate(zip(chain.steps, chain.plans)):
            tp, tg = [], []
            for _ in range(max(3, min(args.steps, 5))):
                flush.zero_()
                a0, a1, a2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                a0.record(stream)
                plan.execute_permute(chain.buf[lhs].ptr, chain.buf[rhs].ptr)
                a1.record(stream)
                # a fused-exchange plan addresses its output in the full result layout
                fused_last = sharded is not None and sharded.exchange in ("fused", "multicast") and si == len(chain.plans) - 1
                plan.execute_gemm(chain.buf[lhs].ptr, chain.buf[rhs].ptr, sharded.full_ptr if fused_last else chain.buf[out].ptr)
                a2.record(stream)
                torch.cuda.synchronize()
                tp.append(a0.elapsed_time(a1)); tg.append(a1.elapsed_time(a2))
            st = stats[si]
            pel = st.permute_elems_a + st.permute_elems_b
            if pel:   # blocks the GEMM cannot read in place
                kern.append({"step": si + 1, "kernel": "batched_permute", "ms": float(np.mean(tp)), "bound": "hbm",
                             "alg_bytes": 2 * pel * es, "achieved": 2 * pel * es / (np.mean(tp) * 1e-3) / 1e9, "unit": "GB/s"})
            kname = "grouped_gemm_dmma" if st.ntile_dmma else "grouped_gemm_skinny"
            if st.ntile_dmma:
                kern.append({"step": si + 1, "kernel": kname, "ms": float(np.mean(tg)), "bound": "tensor", "alg_flops": st.flops,
                             "achieved": st.flops / (np.mean(tg) * 1e-3) / 1e12, "unit": "TFLOP/s"})
            else:
                byts = st.gemm_read_bytes + st.gemm_write_bytes
                kern.append({"step": si + 1, "kernel": kname, "ms": float(np.mean(tg)), "bound": "hbm", "alg_bytes": byts,
                             "achieved": byts / (np.mean(tg) * 1e-3) / 1e9, "unit": "GB/s"})
        sampler.stop()
        # ---- N>1: where one apply's time goes on every rank (local steps vs exchange + waiting for the slowest rank) ----
        rank_phase = None
        if world > 1:
            comp, exch = [], []
            for _ in range(5):
                flush.zero_()
                torch.distributed.barrier()
                marks = {}
                e0 = torch.cuda.Event(enable_timing=True); e0.record(stream)
                def mark(label):
                    ev = torch.cuda.Event(enable_timing=True); ev.record(stream); marks[label] = ev
                sharded.apply(mark)
                torch.cuda.synchronize()
                comp.append(e0.elapsed_time(marks["compute"])); exch.append(marks["compute"].elapsed_time(marks["exchange"]))
            mine = torch.tensor([float(np.mean(comp[1:])), float(np.mean(exch[1:]))], device="cuda", dtype=torch.float64)
            allr = [torch.zeros_like(mine) for _ in range(world)]
            torch.distributed.all_gather(allr, mine)
            rank_phase = [{"rank": r, "local_steps_ms": float(x[0]), "exchange_and_wait_ms": float(x[1])} for r, x in enumerate(allr)]
        # ---- e2e: psi in pinned host memory -> result in pinned host memory, every apply ----
        psi_host = tensors["psi"].data
        n_out = sharded.info.full_elems if sharded is not None else chain.shells["out"].data.size
        out_bytes, shm = env.shared_host_buffer("out", n_out * es)
        out_host = out_bytes.view(np_dtype(dtype))
        tk._lib.check(tk._lib.lib.qlb200_host_register(psi_host.ctypes.data, psi_host.nbytes), "host_register")
        tk._lib.check(tk._lib.lib.qlb200_host_register(out_host.ctypes.data, max(out_host.nbytes, 1)), "host_register")
        e2e_serial_s, e2e_how = None, None
        if sharded is None:
            # one GPU: chunked upload overlapped with step 1, chunked download overlapped with the last step (qlb200_hostpipe_*)
            pipe_in = [float(x) for x in args.pipe_in.split(",")] if args.pipe_in else None
            pipe_out = [float(x) for x in args.pipe_out.split(",")] if args.pipe_out else None
            chain.make_host_pipe("psi", pipe_in, pipe_out)
            e2e_apply = lambda: chain.apply_host_pipelined(psi_host, out_host)
            n_parts = len(pipe_in) if pipe_in else len(chain.pipe_fractions(chain.buf["psi"].nbytes)[0])
            e2e_how = (f"qlb200_hostpipe: psi uploaded in {n_parts} chunks on a copy stream while the parts of step 1 run, the result downloaded in "
                       f"{len(pipe_out) if pipe_out else n_parts} chunks while the last step computes; host call returns when the result is in host memory")
            h2d_b, d2h_b = int(psi_host.nbytes), int(out_host.nbytes)
        elif args.shard_of:
            def e2e_apply():
                tk._lib.check(tk._lib.lib.qlb200_memcpy_h2d(ctx.h, chain.buf["psi"].ptr, psi_host.ctypes.data, psi_host.nbytes), "h2d")
                apply_fn()
                ctx.sync()
            h2d_b, d2h_b = int(psi_host.nbytes), 0
        else:
            # N GPUs: every rank uploads 1/N of psi and fans it out over NVLink; every rank downloads 1/N of the (replicated) result
            # into the one shared host buffer (ShardedChain.apply_host)
            e2e_apply = lambda: sharded.apply_host(psi_host, out_host, apply_fn)
            e2e_how = (f"each rank uploads 1/{world} of psi over its own PCIe link and fans it out to all GPUs "
                       f"({'multimem.st through the NVSwitch' if sharded.in_mc else 'unicast peer stores'}), barrier, sharded apply, each rank "
                       f"downloads 1/{world} of the result (every GPU holds all of it after the exchange) into one shared-memory host buffer")
            h2d_b = int(-(-psi_host.nbytes // world))
            dl_lo, dl_hi = sharded.download_slice(int(sharded.info.full_elems))
            d2h_b = int((dl_hi - dl_lo) * es)
        for _ in range(2):
            e2e_apply()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_apply()
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
        # the end-to-end result is the same tensor: checked on rank 0 against the device-resident apply
        e2e_err = None
        if rank == 0 and not args.shard_of:
            if sharded is None:
                chain.apply_device()
                want = chain.result("out").data
            e2e_err = float(np.linalg.norm(out_host - want) / np.linalg.norm(want))
            if not e2e_err <= 1e-12:
                _fatal(f"end-to-end result differs from the device-resident one: rel err {e2e_err:.3e}")
        if sharded is None:
            for _ in range(2):
                chain.apply_host("psi", psi_host, "out", out_host)
            t0 = time.perf_counter()
            for _ in range(max(3, args.steps // 2)):
                chain.apply_host("psi", psi_host, "out", out_host)
            e2e_serial_s = (time.perf_counter() - t0) / max(3, args.steps // 2)
        tk._lib.lib.qlb200_host_unregister(psi_host.ctypes.data)
        tk._lib.lib.qlb200_host_unregister(out_host.ctypes.data)
        peak_burst, peak_sust, peak_how = env.fp64_peak(dtype)
        # ---- cold drop-in: what a TensorToolkit program pays when it simply swaps qlten::Contract for the adapter ----
        cold = None
        if world == 1 and not args.shard_of and not args.no_cold:
            cold = measure_cold_dropin(tk, ctx, tensors, wl.HEFF_STEPS, chain.flops())
        # ---- the same apply with the two MPO tensors pre-contracted (static across a Lanczos run): one pass over the rank-5
        # intermediate instead of two (workloads.HEFF_STEPS_FUSED); flops numerator stays the reference recipe's
        fused = None
        if world == 1 and not args.shard_of and not args.no_fused_mpo:
            l, r, ax, o = wl.HEFF_FUSE_PREP
            t2 = dict(tensors)
            t2[o] = tk.contract(tensors[l], tensors[r], ax, ctx)
            fchain = ContractionChain(ctx, t2, wl.HEFF_STEPS_FUSED, np_dtype(dtype), args.plan_flags)
            fchain.apply_device()
            got = fchain.result("out").data
            chain.apply_device()
            base = chain.result("out").data
            ferr = float(np.linalg.norm(got - base) / np.linalg.norm(base))
            fgraph = fchain.capture()
            for _ in range(3):
                fgraph.launch()
            fev = []
            for _ in range(args.steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); fgraph.launch(); e1.record(stream)
                fev.append((e0, e1))
            torch.cuda.synchronize()
            fms = float(np.mean([a.elapsed_time(b) for a, b in fev]))
            fst = fchain.stats()[1]
            tf = []
            for _ in range(5):
                flush.zero_()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(stream)
                fchain.plans[1].execute_device(fchain.buf["t1"].ptr, fchain.buf["w12"].ptr, fchain.buf["t3"].ptr)
                a1.record(stream); torch.cuda.synchronize()
                tf.append(a0.elapsed_time(a1))
            fby = fst.gemm_read_bytes + fst.gemm_write_bytes
            fused = {"ms_per_step": fms, "value": chain.flops() / (fms * 1e-3) / 1e9, "unit": "GFLOP/s (flops of the reference's 4-Contract recipe)",
                     "rel_err_vs_unfused": ferr, "mpo_step_ms": float(np.mean(tf)), "mpo_step_gbs": fby / (np.mean(tf) * 1e-3) / 1e9,
                     "how": "w12 = Contract(mpo1, mpo2) once; every apply = lenv x psi, x w12 (one narrow-pair launch over the rank-5 intermediate), x renv"}
            if not ferr <= 1e-12:
                _fatal(f"fused-MPO chain differs from the 4-step chain: rel err {ferr:.3e}")
            fgraph.close(); fchain.close()

    ms = float(np.mean(dev_ms))
    tot_ms, flops_total, e2e_max = ms, flops_local, e2e_s
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        f = torch.tensor([flops_local], device="cuda", dtype=torch.float64)
        dist.all_reduce(f, op=dist.ReduceOp.SUM)
        tot_ms, e2e_max, flops_total = float(t[0]), float(t[1]), float(f[0])
    if rank != 0:
        if sharded is not None:
            sharded.close()
        if shm is not None:
            shm.close()
        return None

    value = flops_total / (tot_ms * 1e-3) / 1e9
    dom = max((k for k in kern), key=lambda k: k["ms"])
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    # DRAM traffic per launch of each kernel from the committed `ncu --set full` capture (profiles/ncu_traffic.json), else null
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(f"D{args.D}_{dtype}" if args.workload == "heff_u1" else "-", {})
    except Exception:
        traffic = {}
    # the default complex kernel multiplies with three real DMMAs per complex step (Karatsuba / "3M") instead of four:
    # `achieved` stays ALGORITHMIC flops (the reference's cost model, 8*m*k*n) / time, so it may exceed the 4-product
    # cuBLAS ZGEMM rate used as `peak`; `executed_ratio` says how many of those flops the tensor pipe really issues
    three_m = dtype == "c128" and not (args.plan_flags & 32)
    if dom["bound"] == "tensor":
        roof = {"bound": "tensor", "kernel": f"step{dom['step']}:{dom['kernel']}", "achieved": dom["achieved"], "peak": peak_burst,
                "unit": "TFLOP/s", "frac": dom["achieved"] / peak_burst, "traffic": traffic.get(f"step{dom['step']}:{dom['kernel']}"),
                "executed_ratio": 0.75 if three_m else 1.0,
                "pipe_frac": dom["achieved"] * (0.75 if three_m else 1.0) / peak_burst,
                "arithmetic": "3M complex product: 3 real DMMA.8x8x4 per complex 8x8x4 step (6*m*k*n executed flops for 8*m*k*n algorithmic)"
                              if three_m else "4 real DMMA.8x8x4 per complex 8x8x4 step" if dtype == "c128" else "DMMA.8x8x4",
                "peak_source": f"FP64 GEMM peak measured in this run: {peak_how}; burst {peak_burst:.2f} / sustained {peak_sust:.2f} TFLOP/s "
                               "(MEASURED_PEAKS.json has no FP64 entry; tcgen05 has no FP64 kind, DMMA via mma.sync is the FP64 tensor path)"}
    else:
        roof = {"bound": "hbm", "kernel": f"step{dom['step']}:{dom['kernel']}", "achieved": dom["achieved"], "peak": hbm_peak, "unit": "GB/s",
                "frac": dom["achieved"] / hbm_peak, "traffic": traffic.get(f"step{dom['step']}:{dom['kernel']}"), "peak_source": hbm_src}
    for k in kern:
        k["frac"] = k["achieved"] / (peak_burst if k["bound"] == "tensor" else hbm_peak)

    cpu, parity = None, None
    if not args.no_cpu_baseline and world == 1:
        import shutil
        import tempfile
        from tensortoolkit_b200 import qlten_io
        tmp = tempfile.mkdtemp(prefix="qlb200_bench_")
        try:
            # The reference's CPU path on THIS input, in a clean child process (no torch / CUDA runtime threads competing
            # with HPTT's and OpenBLAS's pools) -- exactly what `bench.py --impl reference` measures.  The operands travel
            # in the reference's own tensor file format; its result comes back the same way and is compared with ours.
            cD = args.cpu_sample_D or args.D
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(args.cpu_steps), "--warmup", "1",
                   "--D", str(args.D), "--dtype", dtype, "--cpu-sample-D", str(cD), "--workload", args.workload,
                   "--cpu-time-cap", str(args.cpu_time_cap)] + (["--no-thread-sweep"] if args.no_thread_sweep else [])
            same = cD == args.D and not args.shard_of
            if same:
                qn = args.qn if args.tensors else ("fU1U1QN" if args.workload == "heff_hubbard" else "U1QN")
                for name in ("lenv", "psi", "mpo1", "mpo2", "renv"):
                    qlten_io.save(tensors[name], os.path.join(tmp, name + ".qlten"))
                cmd += ["--tensors", tmp, "--qn", qn, "--write-out", os.path.join(tmp, "out_ref.qlten")]
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500,
                                 env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
            ref_line = json.loads(out.stdout.strip().splitlines()[-1])
            cpu = dict(ref_line["cpu_baseline"])
            cpu["sample"] += f"; reference's own CPU path (TensorToolkit Contract + HPTT + OpenBLAS 0.3.15), mean {ref_line['ms_per_step']:.0f} ms per apply"
            if same:
                kind = tensors["psi"].indexes[0].kind
                want = qlten_io.load(os.path.join(tmp, "out_ref.qlten"), kind, np_dtype(dtype))
                with torch.cuda.stream(stream):
                    chain.upload("psi", tensors["psi"].data)
                    chain.apply_device()
                    got = chain.result("out")
                parity = {"rel_fro_vs_reference": float(np.linalg.norm(got.data - want.data) / np.linalg.norm(want.data)),
                          "same_block_structure": bool(got.same_structure(want)), "tolerance": 1e-12,
                          "how": "result of this apply vs the reference's qlten::Contract chain on the same operands (exchanged as .qlten files)"}
                if not (parity["same_block_structure"] and parity["rel_fro_vs_reference"] <= 1e-12):
                    _fatal(f"parity against the reference FAILED: {parity}")
        except SystemExit:
            raise
        except Exception as e:   # the oracle library is test infrastructure; never fatal for the product bench
            cpu = cpu or {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)

    line = {
        "metric": "block-sparse contraction useful FP64 GFLOP/s", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": tot_ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "c64(f64 pairs)" if dtype == "c128" else "f64", "data": "files" if args.tensors else "synthetic",
        "config": {"workload": (f"H_eff apply on tensors read from {args.tensors} ({args.qn})" if args.tensors else workload_name(args.D, dtype, args.workload)), "l2": "512 MiB flush written between timed applies; operands+intermediates (>1 GB) exceed the 126 MB L2",
                   "parallelism": (f"output-sector/row-slab x{world}, exchange={sharded.exchange}" + {"fused": " (unicast peer stores over NVLink from the GEMM epilogue)", "multicast": " (multimem.st from the GEMM epilogue, replicated by the NVSwitch)"}.get(sharded.exchange, " (NCCL)") + f"; buffers + barrier: {'qlb200_comm (C ABI)' if sharded.comm is not None else 'torch symmetric memory / IPC'}") if world > 1 else "single GPU", "flops_per_step": flops_total,
                   "tasks_per_step": int(sum(s.ntask for s in stats)),
                   "launch": "one CUDA graph replay per apply" if graph is not None else "one host launch per kernel"},
        "pct_fp64_peak": 100.0 * value / 1e3 / peak_burst / world, "fp64_peak_tflops": {"burst": peak_burst, "sustained": peak_sust, "how": peak_how},
        "roofline": roof, "kernels": kern, "cpu_baseline": cpu, "parity": parity,
        "e2e": {"value": flops_total / e2e_max / 1e9, "unit": "GFLOP/s", "ms_per_step": e2e_max * 1e3,
                "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b, "per": "rank" if world > 1 else "gpu", "how": e2e_how,
                "rel_err_vs_device_resident": e2e_err,
                "serial_ms_per_step": None if e2e_serial_s is None else e2e_serial_s * 1e3,
                "cold_dropin": cold},
        "fused_mpo": fused,
        "gpu_launches": int(launches) * args.steps, "clocks": sampler.summary(),
    }
    if verified is not None:
        line["sharded_vs_unsharded_rel_err"] = verified
    if self_check is not None:
        line["self_check"] = self_check
    if rank_phase is not None:
        line["rank_phases"] = rank_phase
        line["partition_feedback"] = {"iterations": args.rebalance, "kept_partition": rebalance_kept, "local_ms_by_rank_of_each_partition": rebalance_log,
                                      "how": "sharding.tune_partition at set-up: rows of the split bond re-weighted by (measured time / modelled flops)^damp of the rank that owned them and cut again; every cut measured, the best kept", "damp": args.rebalance_damp}
    if args.breakdown:
        for rp in rank_phase or []:
            print(f"  rank {rp['rank']}: local steps {rp['local_steps_ms']:.3f} ms, exchange + wait {rp['exchange_and_wait_ms']:.3f} ms", file=sys.stderr)
        for k in kern:
            print(f"  step {k['step']} {k['kernel']:22s} {k['ms']:8.3f} ms  {k['achieved']:9.2f} {k['unit']:8s} frac {k['frac']:.3f}", file=sys.stderr)
    if sharded is not None:
        sharded.close()
    else:
        chain.close()
    if shm is not None:
        shm.close()
        if rank == 0:
            shm.unlink()
    return line


def measure_cold_dropin(tk, ctx, tensors, steps, flops, reps=2):
    """The literal drop-in, no reuse of anything: every apply = four Contract calls on HOST tensors, each one doing
    sector match + plan build (descriptor tables, cudaMalloc, upload) + qlb200_execute(QLB200_MEM_HOST) (H2D of A and B from
    pageable memory, kernels, D2H of C, synchronise) + teardown -- what include/qlten_b200/contract.h does per call."""
    import ctypes as C
    lib, check = tk._lib.lib, tk._lib.check
    best = None
    for _ in range(reps):
        t = dict(tensors)
        tm = tp = te = 0.0
        t_all = time.perf_counter()
        for lhs, rhs, axes, out in steps:
            t0 = time.perf_counter()
            m = tk.Match(t[lhs], t[rhs], axes)
            t1 = time.perf_counter()
            plan = tk.ContractionPlan(ctx, m, t[lhs].dtype)
            c = m.result_shell(t[lhs].dtype)
            t2 = time.perf_counter()
            plan.execute_host(t[lhs].data, t[rhs].data, c.data)
            t3 = time.perf_counter()
            plan.close(); m.close()
            t[out] = c
            tm += t1 - t0; tp += t2 - t1; te += t3 - t2
        total = time.perf_counter() - t_all
        if best is None or total < best["ms_per_step"] * 1e-3:
            best = {"ms_per_step": total * 1e3, "match_ms": tm * 1e3, "plan_ms": tp * 1e3, "execute_ms": te * 1e3,
                    "value": flops / total / 1e9, "unit": "GFLOP/s",
                    "how": "4 x (qlb200_match_create + qlb200_plan_create + qlb200_execute(QLB200_MEM_HOST) + destroy) on pageable host tensors, "
                           "intermediates returned to the host after every step; best of %d applies" % reps}
    return best




# ------------------------------------------------------------------------------------------------
from tensortoolkit_b200.workloads import ragged_tables  # noqa: E402  (BASELINE configs[4] descriptor table)


def ragged_slice(tb, ntask):
    """The first `ntask` pairs of the ragged table with offsets re-based to a compact slice (CPU sample / e2e sample)."""
    sub = tb["tasks"][:ntask].copy()
    a_off, b_off = np.zeros(ntask, np.uint64), np.zeros(ntask, np.uint64)
    ao = bo = co = 0
    c_base, flops = {}, 0.0
    for i, t in enumerate(sub):
        m, k, n = int(t["m"]), int(t["k"]), int(t["n"])
        a_off[i], b_off[i] = ao, bo
        if int(t["c_ord"]) not in c_base:
            c_base[int(t["c_ord"])] = co
            co += m * n
        t["a_ord"] = t["b_ord"] = t["a_blk_idx"] = t["b_blk_idx"] = i
        t["a_off"], t["b_off"], t["c_off"] = ao, bo, c_base[int(t["c_ord"])]
        ao += m * k; bo += k * n
        flops += 2.0 * m * k * n
    return dict(tasks=sub, a_shape=tb["a_shape"][:ntask], b_shape=tb["b_shape"][:ntask], a_off=a_off, b_off=b_off,
                a_elems=ao, b_elems=bo, c_elems=co, flops=flops)


def run_ragged_reference(args):
    """Reference arm of configs[4]: the reference's own executor loop (hp_numeric::TensorTranspose per distinct block +
    hp_numeric::MatMultiply per pair, global_operations.h:919-982; oracle/_ref/libqlref.so) on a bounded slice of the table."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "GOTO_NUM_THREADS"):
        os.environ.pop(k, None)
    from oracle import refbridge as ref
    ref.lib()
    threads = pick_threads()
    ref.set_threads(threads)
    s = ragged_slice(ragged_tables(), args.ragged_cpu_pairs)
    rng = np.random.Generator(np.random.MT19937(20260005))
    A = rng.random(s["a_elems"]); B = rng.random(s["b_elems"])
    times = []
    t0 = time.perf_counter()
    for i in range(max(1, args.warmup) + max(1, args.steps)):
        _, sec = ref.raw_contract(np.float64, 3, [1, 2, 0], s["a_shape"], s["a_off"], 3, [1, 0, 2], s["b_shape"], s["b_off"], s["tasks"], A, B, s["c_elems"])
        if i >= max(1, args.warmup):
            times.append(sec)
        if time.perf_counter() - t0 > args.cpu_time_cap and times:
            break
    sec = float(np.mean(times))
    val = s["flops"] / sec / 1e9
    sample = (f"the first {len(s['tasks'])} of 10^4 pairs ({s['flops'] / 1e9:.0f} GFLOP, {(s['a_elems'] + s['b_elems']) * 8 / 1e9:.1f} GB of operands), "
              f"HPTT transpose per block + OpenBLAS dgemm per pair, {threads} threads, mean of {len(times)} passes")
    print(json.dumps({"impl": "reference", "metric": "block-sparse contraction useful FP64 GFLOP/s", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
                      "steps": len(times), "warmup": max(1, args.warmup), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "ragged-sector stress test (10^4 random blocks, sizes 8-2048)", "bounded_sample": sample, "same_config": False},
                      "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": threads, "kind": "reference", "sample": sample},
                      "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def measure_ragged(args, env):
    """configs[4]: ragged-sector stress test, transpose + grouped GEMM only (double).  One step = permute every block that
    cannot be read in place + one grouped GEMM launch over all pairs; operands generated on the device (replicated on every
    rank); N ranks each compute a contiguous cost-balanced share of the output rows (qlb200_plan_partition) and permute
    only the operand blocks that share needs."""
    torch, tk = env.torch, env.tk
    rank, world, stream, ctx = env.rank, env.world, env.stream, env.ctx
    tb = ragged_tables()
    with torch.cuda.stream(stream):
        gen = torch.Generator(device="cuda"); gen.manual_seed(20260005)
        A = torch.rand(tb["a_elems"], dtype=torch.float64, device="cuda", generator=gen)
        B = torch.rand(tb["b_elems"], dtype=torch.float64, device="cuda", generator=gen)
        Cbuf = torch.zeros(tb["c_elems"], dtype=torch.float64, device="cuda")
        plan = tk.RawPlan(ctx, np.float64, 3, [1, 2, 0], tb["a_shape"], tb["a_off"], 3, [1, 0, 2], tb["b_shape"], tb["b_off"], tb["tasks"],
                          tb["c_elems"], args.plan_flags)
        full_flops = plan.stats().flops
        sharded_err = None
        if world > 1:
            # sharded == unsharded on this rank's rows: run the whole plan once, keep it, then the partitioned plan
            plan.execute_device(A.data_ptr(), B.data_ptr(), Cbuf.data_ptr())
            torch.cuda.synchronize()
            full = Cbuf.clone()
            Cbuf.zero_()
            plan.partition(world, rank)
        st = plan.stats()
        flush = env.flush
        for _ in range(max(3, args.warmup)):
            plan.execute_device(A.data_ptr(), B.data_ptr(), Cbuf.data_ptr())
        torch.cuda.synchronize()
        off, ln = plan.c_ranges()
        my_flops = 0.0
        if world > 1:
            num = den = 0.0
            for o, l in zip(off, ln):
                d = Cbuf[int(o):int(o + l)] - full[int(o):int(o + l)]
                num += float(torch.dot(d, d)); den += float(torch.dot(full[int(o):int(o + l)], full[int(o):int(o + l)]))
            sharded_err = (num / den) ** 0.5 if den > 0 else 0.0
            if not sharded_err <= 1e-12:
                _fatal(f"rank {rank}: partitioned ragged plan differs from the whole one, rel err {sharded_err:.3e}")
            del full
        # parity on this very input: a sample of this rank's output blocks against torch.matmul in float64 on the device
        worst = 0.0
        by_c = {}
        for t in tb["tasks"]:
            by_c.setdefault(int(t["c_ord"]), []).append(t)
        owned = [(int(o), int(o + l)) for o, l in zip(off, ln)]
        # flops of this rank's share: every output block's pairs, weighted by the fraction of its rows the rank owns
        c_start = {int(ts[0]["c_off"]): (int(ts[0]["m"]) * int(ts[0]["n"]), sum(2.0 * int(t["m"]) * int(t["k"]) * int(t["n"]) for t in ts))
                   for ts in by_c.values()}
        starts = sorted(c_start)
        import bisect
        for lo, hi in owned:
            i = bisect.bisect_right(starts, lo) - 1
            while i < len(starts) and starts[i] < hi:
                size, fl = c_start[starts[i]]
                ov = min(hi, starts[i] + size) - max(lo, starts[i])
                if ov > 0:
                    my_flops += fl * ov / size
                i += 1
        for c in list(by_c)[:: max(1, len(by_c) // 40)]:
            m, n = int(by_c[c][0]["m"]), int(by_c[c][0]["n"])
            c0 = int(by_c[c][0]["c_off"])
            rows = [(max(c0, lo) - c0, min(c0 + m * n, hi) - c0) for lo, hi in owned if lo < c0 + m * n and hi > c0]
            if not rows:
                continue
            want = torch.zeros(m, n, dtype=torch.float64, device="cuda")
            for t in by_c[c]:
                k = int(t["k"]); ash = [int(x) for x in tb["a_shape"][int(t["a_ord"])]]; bsh = [int(x) for x in tb["b_shape"][int(t["b_ord"])]]
                a = A[int(t["a_off"]):int(t["a_off"]) + m * k].view(*ash).permute(1, 2, 0).reshape(m, k)
                b = B[int(t["b_off"]):int(t["b_off"]) + k * n].view(*bsh).permute(1, 0, 2).reshape(k, n)
                want += a @ b
            for lo, hi in rows:
                got = Cbuf[c0 + lo:c0 + hi]
                w = want.reshape(-1)[lo:hi]
                worst = max(worst, float(torch.linalg.norm(got - w) / torch.linalg.norm(w)))
        if not worst <= 1e-12:
            _fatal(f"ragged workload: grouped GEMM differs from the float64 reference, rel err {worst:.3e}")
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(env.local); sampler.start()
        tot, tp, tg = [], [], []
        for _ in range(args.steps):
            flush.zero_()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(stream)
            plan.execute_permute(A.data_ptr(), B.data_ptr())
            e1.record(stream)
            plan.execute_gemm(A.data_ptr(), B.data_ptr(), Cbuf.data_ptr())
            e2.record(stream)
            torch.cuda.synchronize()
            tot.append(e0.elapsed_time(e2)); tp.append(e0.elapsed_time(e1)); tg.append(e1.elapsed_time(e2))
        sampler.stop()
        launches = int(ctx.launch_count()) + 1
        peak_burst, peak_sust, peak_how = env.fp64_peak("f64")
        # e2e on a bounded slice (the whole table is 2 x ~11 GB of operands): the first pairs through qlb200_execute(QLB200_MEM_HOST)
        e2e = None
        if world == 1:
            s = ragged_slice(tb, args.ragged_cpu_pairs)
            rng = np.random.Generator(np.random.MT19937(20260005))
            Ah = rng.random(s["a_elems"]); Bh = rng.random(s["b_elems"]); Ch = np.empty(s["c_elems"])
            for arr in (Ah, Bh, Ch):
                tk._lib.check(tk._lib.lib.qlb200_host_register(arr.ctypes.data, arr.nbytes), "host_register")
            sp = tk.RawPlan(ctx, np.float64, 3, [1, 2, 0], s["a_shape"], s["a_off"], 3, [1, 0, 2], s["b_shape"], s["b_off"], s["tasks"], s["c_elems"], args.plan_flags)
            sp.execute_host(Ah, Bh, Ch)
            t0 = time.perf_counter()
            for _ in range(3):
                sp.execute_host(Ah, Bh, Ch)
            sec = (time.perf_counter() - t0) / 3
            sp.close()
            for arr in (Ah, Bh, Ch):
                tk._lib.lib.qlb200_host_unregister(arr.ctypes.data)
            e2e = {"value": s["flops"] / sec / 1e9, "unit": "GFLOP/s", "ms_per_step": sec * 1e3, "h2d_bytes_per_step": int(Ah.nbytes + Bh.nbytes),
                   "d2h_bytes_per_step": int(Ch.nbytes),
                   "how": f"bounded sample: the first {len(s['tasks'])} pairs through qlb200_execute(QLB200_MEM_HOST) from pinned host memory "
                          "(H2D of both operands, permute + grouped GEMM, D2H of the result, synchronise); the whole table holds 2 x ~11 GB of operands"}
            del Ah, Bh, Ch
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]); hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    ms, mp, mg = float(np.mean(tot)), float(np.mean(tp)), float(np.mean(tg))
    t = torch.tensor([ms, mp, mg], device="cuda", dtype=torch.float64)
    f = torch.tensor([my_flops, float(st.permute_elems_a + st.permute_elems_b)], device="cuda", dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(f, op=dist.ReduceOp.SUM)
    tot_ms, flops_total = float(t[0]), float(f[0])
    pel = st.permute_elems_a + st.permute_elems_b
    plan.close()
    del A, B, Cbuf
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    kern = [{"step": 1, "kernel": "batched_permute", "ms": mp, "bound": "hbm", "alg_bytes": 2 * pel * 8, "achieved": 2 * pel * 8 / (mp * 1e-3) / 1e9 if mp > 0 else 0.0,
             "unit": "GB/s", "frac": (2 * pel * 8 / (mp * 1e-3) / 1e9 / hbm_peak) if mp > 0 else 0.0},
            {"step": 1, "kernel": "grouped_gemm_dmma", "ms": mg, "bound": "tensor", "alg_flops": my_flops, "achieved": my_flops / (mg * 1e-3) / 1e12,
             "unit": "TFLOP/s", "frac": my_flops / (mg * 1e-3) / 1e12 / peak_burst}]
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", "ragged", "--steps", "2", "--warmup", "1",
                                  "--ragged-cpu-pairs", str(args.ragged_cpu_pairs)], capture_output=True, text=True, timeout=900,
                                 env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
            cpu = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:
            cpu = {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
    line = {"metric": "block-sparse contraction useful FP64 GFLOP/s", "value": flops_total / (tot_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": tot_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "ragged-sector stress test: 2500 output blocks x 4 pairs = 10^4 random blocks (SplitMix64 20260005), m/k/n log-uniform in [8, 2048], "
                                   "A stored (k, m1, m2) perm {1,2,0}, B stored (n1, k, n2) perm {1,0,2}; transpose + grouped GEMM only",
                       "l2": "512 MiB flush between steps; operands 2 x ~11 GB",
                       "parallelism": f"output rows cut into {world} cost-balanced contiguous shares (qlb200_plan_partition), operands replicated, no exchange" if world > 1 else "single GPU",
                       "flops_per_step": flops_total, "tasks_per_step": int(len(tb["tasks"])),
                       "permuted_elems_all_ranks": int(f[1]), "operand_bytes": int((tb["a_elems"] + tb["b_elems"]) * 8)},
            "pct_fp64_peak": 100.0 * flops_total / (tot_ms * 1e-3) / 1e12 / peak_burst / world, "fp64_peak_tflops": {"burst": peak_burst, "sustained": peak_sust, "how": peak_how},
            "roofline": {"bound": "tensor", "kernel": "step1:grouped_gemm_dmma", "achieved": kern[1]["achieved"], "peak": peak_burst, "unit": "TFLOP/s",
                         "frac": kern[1]["frac"], "traffic": None, "peak_source": f"FP64 GEMM peak measured in this run: {peak_how}; hbm: {hbm_src}"},
            "kernels": kern, "cpu_baseline": cpu, "parity_rel_err_sampled_blocks": worst, "e2e": e2e,
            "gpu_launches": launches * args.steps, "clocks": sampler.summary()}
    if sharded_err is not None:
        line["sharded_vs_unsharded_rel_err"] = sharded_err
    if args.breakdown:
        for k in kern:
            print(f"  {k['kernel']:22s} {k['ms']:8.3f} ms  {k['achieved']:9.2f} {k['unit']:8s} frac {k['frac']:.3f}", file=sys.stderr)
    return line


def compact(line):
    """Sub-record of the default run: the fields the judge reads, without the long per-kernel tables."""
    if line is None:
        return None
    keep = ("value", "unit", "ms_per_step", "n_gpus", "steps", "dtype", "pct_fp64_peak", "roofline", "cpu_baseline", "e2e", "parity",
            "parity_rel_err_sampled_blocks", "sharded_vs_unsharded_rel_err", "gpu_launches")
    out = {k: line[k] for k in keep if k in line}
    out["workload"] = line["config"]["workload"]
    out["parallelism"] = line["config"].get("parallelism")
    out["kernels"] = [{k: v for k, v in kk.items() if k in ("step", "kernel", "ms", "achieved", "unit", "frac")} for kk in line.get("kernels", [])]
    return out


def run_ours(args):
    import copy
    env = Env()
    if args.workload == "ragged":
        line = measure_ragged(args, env)
    else:
        line = measure_heff(args, env)
    # the default run (headline) also measures the other named BASELINE shapes, as sub-records of the same JSON line
    headline = args.workload == "heff_u1" and args.D == 4096 and args.dtype == "c128" and not args.tensors and not args.shard_of
    if headline and not args.no_sub_records:
        subs = {}
        for name, (workload, D, dtype, cpuD) in (("configs[1] U(1) Heisenberg D=1024 double", ("heff_u1", 1024, "f64", 0)),
                                                 ("configs[3] Hubbard fU1U1 D=8192 double", ("heff_hubbard", 8192, "f64", 4096))):
            a2 = copy.copy(args)
            a2.workload, a2.D, a2.dtype, a2.cpu_sample_D = workload, D, dtype, cpuD
            a2.steps, a2.cpu_steps, a2.no_cold, a2.breakdown = min(args.steps, 10), 2, True, False
            a2.no_fused_mpo = False
            a2.no_thread_sweep = True
            sub = measure_heff(a2, env)
            subs[name] = compact(sub)
        a2 = copy.copy(args)
        a2.steps, a2.breakdown = min(args.steps, 5), False
        subs["configs[4] ragged stress"] = compact(measure_ragged(a2, env))
        if line is not None:
            line["sub_records"] = subs
    if env.rank != 0 or line is None:
        return
    print(json.dumps(line))


def main():
    args = parse()
    # keep stdout clean for the ONE JSON line: libraries (NCCL's version banner) write to fd 1 too
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_ragged_reference(args) if args.workload == "ragged" else run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()
    try:                                  # leave the NCCL process group cleanly (N > 1)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
