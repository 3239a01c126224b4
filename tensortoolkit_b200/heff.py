"""Device-resident chain of contractions: plan once, apply many times (Lanczos-style mat-vec).

The reference's recipe for the two-site effective-Hamiltonian apply is four chained Contract calls
(tests/test_tensor_manipulation/test_ten_ctrct_1sct.cc:253-257); its block topology is constant
across iterations, so matches, descriptor tables and buffers are built once here and every apply
only enqueues kernels on the context's stream (no host round trips between steps).
"""
import ctypes as C
import os
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import lib, check
from .contract import Context, ContractionPlan, Match
from .tensor import BlockSparseTensor


class DeviceBuffer:
    def __init__(self, ctx: Context, nbytes: int):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(lib.qlb200_dev_alloc(ctx.h, max(self.nbytes, 16), C.byref(p)), "qlb200_dev_alloc")
        self.ptr = p.value

    def upload(self, host: np.ndarray):
        assert host.nbytes <= self.nbytes
        check(lib.qlb200_memcpy_h2d(self.ctx.h, C.c_void_p(self.ptr), host.ctypes.data, host.nbytes), "qlb200_memcpy_h2d")

    def download(self, host: np.ndarray):
        assert host.nbytes <= self.nbytes
        check(lib.qlb200_memcpy_d2h(self.ctx.h, host.ctypes.data, C.c_void_p(self.ptr), host.nbytes), "qlb200_memcpy_d2h")

    def free(self):
        if self.ptr:
            lib.qlb200_dev_free(self.ctx.h, C.c_void_p(self.ptr))
            self.ptr = None


class ExternalBuffer(DeviceBuffer):
    """Device memory owned by the caller."""

    def __init__(self, ctx: Context, ptr: int, nbytes: int):
        self.ctx, self.ptr, self.nbytes = ctx, int(ptr), int(nbytes)

    def free(self):
        self.ptr = None


class CapturedGraph:
    """A CUDA graph of everything `fn` enqueued on the context's stream (qlb200_graph_*): one launch per replay."""

    def __init__(self, ctx: Context, fn):
        self.ctx = ctx
        fn()                      # run once eagerly: arenas reach their final size, lazy initialisation is done
        ctx.sync()
        check(lib.qlb200_graph_begin(ctx.h), "qlb200_graph_begin")
        try:
            self.launches = fn()
        finally:
            h = C.c_void_p()
            rc = lib.qlb200_graph_end(ctx.h, C.byref(h))
        check(rc, "qlb200_graph_end")
        self.h = h

    def launch(self):
        check(lib.qlb200_graph_launch(self.ctx.h, self.h), "qlb200_graph_launch")
        return self.launches

    def close(self):
        if self.h:
            lib.qlb200_graph_destroy(self.h)
            self.h = None


class ContractionChain:
    """steps: list of (lhs name, rhs name, axes, out name); `tensors` holds the host operands."""

    def __init__(self, ctx: Context, tensors: Dict[str, BlockSparseTensor], steps: Sequence[Tuple[str, str, tuple, str]],
                 dtype, flags: int = _lib.PLAN_DETERMINISTIC, external: Dict[str, int] = None, last_flags: int = 0):
        self.ctx, self.dtype, self.steps = ctx, np.dtype(dtype), list(steps)
        self.shells: Dict[str, BlockSparseTensor] = dict(tensors)
        self.matches: List[Match] = []
        self.plans: List[ContractionPlan] = []
        for lhs, rhs, axes, out in self.steps:
            m = Match(self.shells[lhs], self.shells[rhs], axes)
            self.matches.append(m)
            self.shells[out] = m.result_shell(self.dtype)
            last = len(self.plans) == len(self.steps) - 1
            self.plans.append(ContractionPlan(ctx, m, self.dtype, flags | (last_flags if last else 0)))
        self.buf: Dict[str, DeviceBuffer] = {}
        for name, t in self.shells.items():
            if external and name in external:      # caller-owned device memory (e.g. a torch tensor)
                self.buf[name] = ExternalBuffer(ctx, external[name], t.data.size * self.dtype.itemsize)
            else:
                self.buf[name] = DeviceBuffer(ctx, t.data.size * self.dtype.itemsize)
        for name, t in tensors.items():
            self.buf[name].upload(t.data)
        ctx.sync()
        self.launches_per_apply = 0

    def stats(self):
        return [p.stats() for p in self.plans]

    def flops(self) -> float:
        return float(sum(s.flops for s in self.stats()))

    def upload(self, name: str, host: np.ndarray):
        self.buf[name].upload(host)

    def apply_device(self):
        """Enqueue all steps; returns the number of kernels launched."""
        n = 0
        for (lhs, rhs, _, out), plan in zip(self.steps, self.plans):
            plan.execute_device(self.buf[lhs].ptr, self.buf[rhs].ptr, self.buf[out].ptr)
            n += self.ctx.launch_count()
        self.launches_per_apply = n
        return n

    def capture(self) -> CapturedGraph:
        """The whole chain as one CUDA graph (replay with .launch())."""
        return CapturedGraph(self.ctx, self.apply_device)

    def apply_host(self, in_name: str, host_in: np.ndarray, out_name: str, host_out: np.ndarray):
        """End-to-end apply: input H2D, all steps, result D2H, synchronise (copies and math strictly in sequence)."""
        self.buf[in_name].upload(host_in)
        self.apply_device()
        self.buf[out_name].download(host_out)
        self.ctx.sync()

    # cumulative shares of the streamed operand per chunk: a small first chunk lets the math start early, a small last
    # output chunk leaves little of the download exposed (the PCIe copies are faster than the steps they overlap)
    # Host pipeline chunking: n parts with cumulative fractions (i/n)^p of the streamed input (small first chunks: step 1 starts
    # early) and 1 - (1 - i/n)^p of the output (small last chunks: little is left to download when the last part ends).
    # Measured at D=4096 complex (exp/r2_call30.sh, r2_call31.sh): 4 parts 9.48 ms, 8 parts 8.84, 12 parts p=2.2 8.78, 24 parts 9.00
    # (device-resident apply 8.26).  Every part costs a launch tail, so small tensors get fewer parts (one per PIPE_BYTES).
    PIPE_PARTS, PIPE_POWER, PIPE_BYTES = 12, 2.2, 8 << 20

    @classmethod
    def pipe_fractions(cls, nbytes: int):
        n = max(1, min(cls.PIPE_PARTS, int(nbytes) // cls.PIPE_BYTES))
        return ([(i / n) ** cls.PIPE_POWER for i in range(1, n + 1)], [1.0 - (1.0 - i / n) ** cls.PIPE_POWER for i in range(1, n + 1)])

    def make_host_pipe(self, in_name: str, cum_in=None, cum_out=None):
        """qlb200_hostpipe over the first and last step: `in_name` (an operand of the first step) is streamed from host
        memory in chunks while the parts of the first step run; the last step's output leaves in chunks while it computes."""
        lhs, rhs, _, _ = self.steps[0]
        if in_name not in (lhs, rhs):
            raise ValueError("the streamed input must be an operand of the first step")
        d_in, d_out = self.pipe_fractions(self.buf[in_name].nbytes)
        cum_in = list(cum_in or d_in); cum_out = list(cum_out or d_out)
        h = C.c_void_p()
        check(lib.qlb200_hostpipe_create(self.ctx.h, self.plans[0].h, _lib.SPLIT_BY_A if in_name == lhs else _lib.SPLIT_BY_B,
                                         len(cum_in), (C.c_double * len(cum_in))(*cum_in), self.plans[-1].h, len(cum_out),
                                         (C.c_double * len(cum_out))(*cum_out), C.byref(h)), "qlb200_hostpipe_create")
        self._pipe, self._pipe_in = h, in_name
        return h

    def apply_host_pipelined(self, host_in: np.ndarray, host_out: np.ndarray) -> int:
        """End-to-end apply through the host pipeline (make_host_pipe first): chunked H2D overlapped with step 1, chunked
        D2H overlapped with the last step.  Returns after the result is in host_out; value = kernels launched."""
        lhs, rhs, _, out = self.steps[0]
        other = rhs if self._pipe_in == lhs else lhs
        check(lib.qlb200_hostpipe_begin(self.ctx.h, self._pipe, host_in.ctypes.data, C.c_void_p(self.buf[self._pipe_in].ptr),
                                        C.c_void_p(self.buf[other].ptr), C.c_void_p(self.buf[out].ptr)), "qlb200_hostpipe_begin")
        n = 0
        for (l, r, _, o), plan in zip(self.steps[1:-1], self.plans[1:-1]):
            plan.execute_device(self.buf[l].ptr, self.buf[r].ptr, self.buf[o].ptr)
            n += self.ctx.launch_count()
        l, r, _, o = self.steps[-1]
        check(lib.qlb200_hostpipe_end(self.ctx.h, self._pipe, C.c_void_p(self.buf[l].ptr), C.c_void_p(self.buf[r].ptr),
                                      C.c_void_p(self.buf[o].ptr), host_out.ctypes.data), "qlb200_hostpipe_end")
        return n + int(lib.qlb200_hostpipe_launches(self._pipe))

    def result(self, name: str) -> BlockSparseTensor:
        t = self.shells[name]
        out = BlockSparseTensor(t.indexes, self.dtype)
        if t.rank:
            out.set_blocks(t.blk_coors)
        else:
            out.data = np.zeros(t.data.size, self.dtype)
        self.buf[name].download(out.data)
        self.ctx.sync()
        return out

    def close(self):
        if getattr(self, "_pipe", None):
            lib.qlb200_hostpipe_destroy(self._pipe)
            self._pipe = None
        for p in self.plans:
            p.close()
        for m in self.matches:
            m.close()
        for b in self.buf.values():
            b.free()
        self.plans, self.matches, self.buf = [], [], {}


class Comm:
    """qlb200_comm: symmetric buffers (unicast peer pointers + NVSwitch multicast mapping) and the device-side barrier, all
    behind the C ABI.  The only thing taken from torch is the bootstrap all-gather of a few bytes (the callback a
    TensorToolkit program would implement with MPI_Allgather)."""

    def __init__(self, ctx: Context, world: int, rank: int, group=None):
        import torch
        import torch.distributed as dist
        self.ctx, self.world, self.rank = ctx, world, rank
        dev = torch.device("cuda", ctx.device)

        def allgather(user, send, recv, nbytes):
            try:
                mine = torch.frombuffer(bytearray(C.string_at(send, nbytes)), dtype=torch.uint8).to(dev)
                outs = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(outs, mine, group=group)
                data = b"".join(bytes(o.cpu().numpy().tobytes()) for o in outs)
                C.memmove(recv, data, len(data))
                return 0
            except Exception as e:      # never let an exception cross the C boundary
                import sys
                print(f"qlb200 all-gather callback failed: {e}", file=sys.stderr)
                return 1
        self._cb = _lib.ALLGATHER_FN(allgather)           # keep alive as long as the communicator
        h = C.c_void_p()
        check(lib.qlb200_comm_create(ctx.h, world, rank, C.cast(self._cb, C.c_void_p) if world > 1 else None, None, C.byref(h)), "qlb200_comm_create")
        self.h = h
        self.has_multicast = bool(lib.qlb200_comm_has_multicast(h))

    def alloc(self, nbytes: int):
        """Collective.  Returns (local pointer, [pointer of every rank's buffer, own first], multicast pointer or 0)."""
        local, mc = C.c_void_p(), C.c_void_p()
        peers = (C.c_void_p * self.world)()
        check(lib.qlb200_comm_alloc(self.h, nbytes, C.byref(local), peers, C.byref(mc)), "qlb200_comm_alloc")
        order = [int(peers[self.rank])] + [int(peers[p]) for p in range(self.world) if p != self.rank]
        return int(local.value), order, int(mc.value or 0)

    def barrier(self):
        check(lib.qlb200_comm_barrier(self.h), "qlb200_comm_barrier")

    def close(self):
        if self.h:
            lib.qlb200_comm_destroy(self.h)
            self.h = None


class _IdleChain:
    """Stand-in for a rank whose share of the split index is empty (world larger than the number of cuttable row
    groups): no plans, no launches -- the rank still takes part in the exchange barrier."""

    def __init__(self, steps):
        self.steps, self.plans, self.matches, self.buf, self.shells = list(steps), [], [], {}, {}
        self.launches_per_apply = 0

    def stats(self):
        return []

    def flops(self) -> float:
        return 0.0

    def apply_device(self):
        return 0

    def close(self):
        pass


class _AlternatingGraphs:
    """The two captured applies of a ShardedChain (one per result replica), launched in turn."""

    def __init__(self, owner, graphs):
        self.owner, self.graphs = owner, graphs
        self.launches = graphs[0].launches

    def launch(self):
        g = self.graphs[self.owner.parity]
        self.owner.parity ^= 1
        return g.launch()

    def close(self):
        for g in self.graphs:
            g.close()


class ShardedChain:
    """One rank's share of a contraction chain plus the exchange that rebuilds the full result on
    every GPU.  Ranks own disjoint output rows (see sharding.py), so no reduction is needed.

    exchange="fused" (default for world > 1): the last step's grouped GEMM stores its output tiles straight
        into the full result buffer of EVERY rank -- its own and, through CUDA-IPC-mapped NVLink peer
        pointers, the others' -- from inside the kernel epilogue (qlb200_execute_bcast); the only collective
        left is a one-element all-reduce that acts as the barrier before the result is read.
    exchange="multicast": as "fused", but the result buffer is symmetric memory (torch.distributed._symmetric_memory)
        with an NVSwitch multicast mapping: every output tile leaves the GPU ONCE as multimem.st stores and the
        switch replicates it into all replicas (qlb200_execute_mcast) -- 1/world of the NVLink traffic of the
        unicast peer stores; the barrier is symmetric memory's signal-pad barrier (no NCCL call per apply).
    exchange="auto": "multicast" when the fabric offers it, else "fused".
    exchange="allgather": local steps -> NCCL all-gather of the packed row slabs -> one batched-copy launch
        that scatters every rank's slabs into the full raw-buffer layout (the library-collective baseline).
    exchange=None: no exchange at all (time one rank's share on a single GPU).

    Iterative use (Lanczos: the result of apply j is the input of apply j+1).  The fused exchanges store into the
    replicas of PEERS, and the only ordering point is the barrier that ends an apply.  With a single result buffer a fast
    rank could reach the exchanged step of apply j+1 and overwrite a slower peer's replica while that peer is still
    reading result j as its input.  The full result is therefore DOUBLE-BUFFERED: apply j writes replica j & 1 on every
    GPU, so the stores of apply j+1 never touch what anybody reads during apply j+1 (result j), and the replica they do
    touch (result j-1) was last read before the barrier that ended apply j.  `full_ptr` is the replica the last apply
    wrote; pass it as the next input."""

    def __init__(self, ctx: Context, tensors: Dict[str, BlockSparseTensor], steps, name: str, axis: int, dtype,
                 world: int, rank: int, group=None, flags: int = _lib.PLAN_DETERMINISTIC, exchange="auto", peers=None,
                 host_input: str = None, pieces=None, snap: int = 8, plumbing: str = "capi"):
        """host_input: name of the operand that arrives from HOST memory every apply (apply_host): its device buffer is
        made exchangeable (symmetric memory / CUDA IPC) so that every rank uploads only 1/world of it."""
        import torch
        from .sharding import shard_chain
        self.torch, self.group = torch, group
        self.ctx, self.world, self.rank = ctx, world, rank
        mine, self.info = shard_chain(tensors, steps, name, axis, world, rank, dtype, pieces=pieces, snap=snap)
        self.dtype = np.dtype(dtype)
        tdt = torch.complex128 if self.dtype == np.complex128 else torch.float64
        dev = torch.device("cuda", ctx.device)
        self.exchange = exchange
        self.stride = max(max(self.info.local_elems), 1)
        self.local = torch.zeros(self.stride, dtype=tdt, device=dev)
        self.out_name = steps[-1][3]
        self.cplan, self.full_buf, self.peer_ptrs, self.opened = None, None, None, []
        # exchanged result: let the last step's output tiles complete throughout its launch, so that the NVLink
        # transfer (every GPU has to RECEIVE the whole result) overlaps the remaining math
        stagger = _lib.PLAN_STAGGER_OUTPUT if (world > 1 and exchange in ("auto", "multicast", "fused")
                                               and os.environ.get("QLB200_STAGGER", "1") != "0") else 0
        # a rank can end up with no rows at all (cuts are snapped to row groups; world may exceed what can be cut):
        # it builds no plans but joins every rendezvous, exchange and barrier below like its peers
        self.idle = self.info.local_elems[rank] == 0
        external = {self.out_name: self.local.data_ptr()}
        self.host_input, self.in_symm, self.in_mc, self.in_peers, self.in_buf = host_input, None, 0, None, None
        # plumbing="capi" (default): symmetric buffers, multicast mapping and barrier come from qlb200_comm_* (no torch
        # symmetric memory, no CUDA IPC, no NCCL call per apply); "torch": the round-1 path kept for comparison
        self.comm = None
        use_capi = plumbing == "capi" and world > 1 and peers is None and exchange in ("auto", "multicast", "fused")
        if use_capi:
            self.comm = Comm(ctx, world, rank, group)
            if exchange == "multicast" and not self.comm.has_multicast:
                raise RuntimeError("exchange='multicast': this NVLink domain offers no multicast (NVLS) mapping")
        if host_input is not None:
            t_in = tensors[host_input]
            self.in_bytes = t_in.data.size * self.dtype.itemsize
            nb = max((self.in_bytes + 255) & ~255, 256)
            if use_capi:
                self.in_ptr, self.in_peers, mc = self.comm.alloc(nb)
                self.in_mc = mc if exchange != "fused" else 0
            elif exchange in ("auto", "multicast") and world > 1 and peers is None:
                import torch.distributed._symmetric_memory as symm_mem
                grp = group if group is not None else torch.distributed.group.WORLD
                with torch.cuda.device(dev):
                    self.in_symm_buf = symm_mem.empty(nb // self.dtype.itemsize, dtype=tdt, device=dev)
                    self.in_symm = symm_mem.rendezvous(self.in_symm_buf, grp)
                self.in_mc = int(self.in_symm.multicast_ptr)
                self.in_ptr = int(self.in_symm_buf.data_ptr())
            else:
                self.in_buf = DeviceBuffer(ctx, nb)
                self.in_ptr = self.in_buf.ptr
            external[host_input] = self.in_ptr
        if self.idle:
            self.chain = _IdleChain(steps)
        else:
            self.chain = ContractionChain(ctx, mine, steps, dtype, flags, external=external, last_flags=stagger)
        full_bytes = max(self.info.full_elems, 1) * self.dtype.itemsize
        self.full_bytes = (full_bytes + 255) & ~255          # replica stride (keeps both halves 256-byte aligned)
        self.parity = 0                                      # replica the NEXT apply writes
        self.full_base = None
        self.symm = None
        if use_capi:
            self.full_base, order, mc = self.comm.alloc(2 * self.full_bytes)
            my = self.info.slabs[rank]
            if not self.idle:
                self.chain.plans[-1].remap_output([s.local_offset for s in my], [s.full_offset for s in my])
            if mc and exchange != "fused":
                exchange, self.mc_base = "multicast", mc
            else:
                exchange, self.peer_ptrs = "fused", order
                self.flag = None
        elif exchange in ("auto", "multicast") and world > 1 and peers is None:
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else torch.distributed.group.WORLD
            with torch.cuda.device(dev):
                self.symm_buf = symm_mem.empty(2 * self.full_bytes // self.dtype.itemsize, dtype=tdt, device=dev)
                self.symm = symm_mem.rendezvous(self.symm_buf, grp)
            if not int(self.symm.multicast_ptr):
                if exchange == "multicast":
                    raise RuntimeError("exchange='multicast': this NVLink domain offers no multicast (NVLS) mapping")
                self.symm, self.symm_buf, exchange = None, None, "fused"
            else:
                exchange = "multicast"
                self.full_base = int(self.symm_buf.data_ptr())
                self.mc_base = int(self.symm.multicast_ptr)
                my = self.info.slabs[rank]
                if not self.idle:
                    self.chain.plans[-1].remap_output([s.local_offset for s in my], [s.full_offset for s in my])
        elif exchange == "auto":
            exchange = "fused"
        self.exchange = exchange
        if exchange == "multicast" or use_capi:
            pass
        elif exchange == "fused":
            # the full result lives in a cudaMalloc'ed buffer of its own so that it can be exported over CUDA IPC
            self.full_buf = DeviceBuffer(ctx, 2 * self.full_bytes)
            self.full_base = self.full_buf.ptr
            my = self.info.slabs[rank]
            if not self.idle:
                self.chain.plans[-1].remap_output([s.local_offset for s in my], [s.full_offset for s in my])
            if peers is not None:                       # caller-supplied replicas (single-process tests)
                self.peer_ptrs = [self.full_base] + [int(x) for x in peers]      # each 2 * full_bytes long
            elif world > 1:
                handle = C.create_string_buffer(64)
                check(lib.qlb200_ipc_export(ctx.h, C.c_void_p(self.full_base), handle), "qlb200_ipc_export")
                handles = [None] * world
                torch.distributed.all_gather_object(handles, bytes(handle.raw), group=group)
                self.peer_ptrs = [self.full_base]
                for r in range(world):
                    if r == rank:
                        continue
                    pp = C.c_void_p()
                    check(lib.qlb200_ipc_open(ctx.h, handles[r], C.byref(pp)), "qlb200_ipc_open")
                    self.opened.append(pp.value)
                    self.peer_ptrs.append(pp.value)
            else:
                self.peer_ptrs = [self.full_base]
            self.flag = torch.zeros(1, dtype=torch.float32, device=dev)
            if host_input is not None and world > 1 and peers is None and not self.in_mc:
                handle = C.create_string_buffer(64)       # the input buffer, exchanged like the result
                check(lib.qlb200_ipc_export(ctx.h, C.c_void_p(self.in_ptr), handle), "qlb200_ipc_export")
                handles = [None] * world
                torch.distributed.all_gather_object(handles, bytes(handle.raw), group=group)
                self.in_peers = [self.in_ptr]
                for r in range(world):
                    if r != rank:
                        pp = C.c_void_p()
                        check(lib.qlb200_ipc_open(ctx.h, handles[r], C.byref(pp)), "qlb200_ipc_open")
                        self.opened.append(pp.value)
                        self.in_peers.append(pp.value)
        else:
            self.gathered = torch.zeros(world * self.stride, dtype=tdt, device=dev)
            self.full2 = torch.zeros(2 * self.full_bytes // self.dtype.itemsize, dtype=tdt, device=dev)
            self.full_base = self.full2.data_ptr()
            src, dst, ln = [], [], []
            for r in range(world):
                for s in self.info.slabs[r]:
                    src.append(r * self.stride + s.local_offset); dst.append(s.full_offset); ln.append(s.length)
            n = len(src)
            arr = lambda v: (C.c_uint64 * max(n, 1))(*v)
            h = C.c_void_p()
            check(lib.qlb200_cplan_create(ctx.h, _lib.C64 if self.dtype == np.complex128 else _lib.F64, n, arr(src), arr(dst), arr(ln),
                                          C.byref(h)), "qlb200_cplan_create")
            self.cplan = h

    def flops_local(self) -> float:
        return self.chain.flops()

    def apply(self, mark=None) -> int:
        """All steps of this rank plus the exchange, enqueued on the current torch stream (== ctx stream).
        `mark(label)` (optional) is called after the local steps and after the exchange (timing hooks)."""
        ch = self.chain
        mark = mark or (lambda label: None)
        off = self.parity * self.full_bytes       # this apply's replica
        self.parity ^= 1
        if self.exchange == "multicast":
            n = 0
            for (lhs, rhs, _, out), plan in zip(ch.steps[:-1], ch.plans[:-1]):
                plan.execute_device(ch.buf[lhs].ptr, ch.buf[rhs].ptr, ch.buf[out].ptr)
                n += self.ctx.launch_count()
            if not self.idle:
                lhs, rhs, _, _ = ch.steps[-1]
                ch.plans[-1].execute_mcast(ch.buf[lhs].ptr, ch.buf[rhs].ptr, self.mc_base + off)
                n += self.ctx.launch_count()
            mark("compute")
            self._barrier()          # every rank's tiles have landed in every replica
            mark("exchange")
            return n
        if self.exchange == "fused":
            n = 0
            for (lhs, rhs, _, out), plan in zip(ch.steps[:-1], ch.plans[:-1]):
                plan.execute_device(ch.buf[lhs].ptr, ch.buf[rhs].ptr, ch.buf[out].ptr)
                n += self.ctx.launch_count()
            if not self.idle:
                lhs, rhs, _, _ = ch.steps[-1]
                ch.plans[-1].execute_bcast(ch.buf[lhs].ptr, ch.buf[rhs].ptr, [p + off for p in self.peer_ptrs])
                n += self.ctx.launch_count()
            mark("compute")
            if self.world > 1 and (self.opened or self.comm is not None):
                self._barrier()      # every peer's tiles have landed
            mark("exchange")
            return n
        n = ch.apply_device()
        mark("compute")
        if self.exchange is None:
            return n
        if self.world > 1:
            self.torch.distributed.all_gather_into_tensor(self.gathered, self.local, group=self.group)
        else:
            self.gathered.copy_(self.local)
        check(lib.qlb200_copy_execute(self.ctx.h, self.cplan, C.c_void_p(self.gathered.data_ptr()), C.c_void_p(self.full_base + off)),
              "qlb200_copy_execute")
        mark("exchange")
        return n + 1

    def _barrier(self):
        """Barrier across the ranks on the context's stream: qlb200_comm_barrier (device-side epochs through peer memory),
        or with plumbing='torch' the symmetric-memory signal-pad barrier / a one-element NCCL all-reduce."""
        if self.comm is not None:
            self.comm.barrier()
        elif self.symm is not None:
            self.symm.barrier()
        elif self.in_symm is not None:
            self.in_symm.barrier()
        else:
            self.torch.distributed.all_reduce(self.flag, group=self.group)

    def barrier(self):
        """Public form of the stream-ordered barrier across the ranks (e.g. to let all ranks start a timed apply together)."""
        if self.world > 1:
            self._barrier()

    @property
    def full_ptr(self) -> int:
        """Device address of the full result the LAST apply produced (before the first apply: replica 0)."""
        return self.full_base + (self.parity ^ 1) * self.full_bytes if self.full_base else 0

    @property
    def full(self):
        """torch view of the full result the last apply produced (exchange='allgather' only)."""
        n = max(self.info.full_elems, 1)
        o = (self.parity ^ 1) * self.full_bytes // self.dtype.itemsize
        return self.full2[o:o + n]

    def capture(self):
        """Local steps + exchange + barrier as CUDA graphs, one per result replica, replayed alternately (the torch
        stream must be the context's stream)."""
        graphs = []
        for par in (0, 1):
            def fn(par=par):
                self.parity = par
                return self.apply()
            graphs.append(CapturedGraph(self.ctx, fn))
        self.parity = 0
        return _AlternatingGraphs(self, graphs)

    def own_ranges(self):
        """Merged (element offset, length) ranges of the full result this rank computes."""
        out = []
        for s in self.info.slabs[self.rank]:
            if out and out[-1][0] + out[-1][1] == s.full_offset:
                out[-1][1] += s.length
            else:
                out.append([s.full_offset, s.length])
        return out

    def apply_host(self, host_in: np.ndarray, host_out: np.ndarray, apply_fn=None) -> int:
        """End-to-end apply with the input in (pinned) host memory and the result back in host memory, N ranks:
        every rank uploads only ITS 1/world share of the input over its own PCIe link and fans it out to all GPUs
        (multimem.st through the NVSwitch, or unicast peer stores); after one barrier the whole input is everywhere.
        Then the sharded apply (`apply_fn`, default self.apply; a captured graph's launch works too) and the download of
        this rank's equal slice of the full result (download_slice) into host_out at its place in the full layout -- the
        ranks' downloads together are the full result (host_out may be one shared-memory buffer mapped by all ranks)."""
        es = self.dtype.itemsize
        if self.host_input is None:
            raise RuntimeError("ShardedChain was built without host_input")
        world, rank = self.world, self.rank
        if world > 1 and (self.in_mc or self.in_peers):
            unit = 16 // es if es < 16 else 1
            n = host_in.size
            per = -(-n // world)
            per = -(-per // unit) * unit
            lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
            if hi > lo:
                check(lib.qlb200_memcpy_h2d(self.ctx.h, C.c_void_p(self.in_ptr + lo * es), host_in.ctypes.data + lo * es, (hi - lo) * es), "h2d")
                nbytes = ((hi - lo) * es + 15) & ~15      # the buffer is padded to 256 bytes: rounding the tail up is safe
                if self.in_mc:
                    check(lib.qlb200_fanout_copy(self.ctx.h, C.c_void_p(self.in_ptr), lo * es, nbytes, None, 0, C.c_void_p(self.in_mc)), "fanout")
                else:
                    arr = (C.c_void_p * len(self.in_peers))(*self.in_peers)
                    check(lib.qlb200_fanout_copy(self.ctx.h, C.c_void_p(self.in_ptr), lo * es, nbytes, arr, len(self.in_peers), None), "fanout")
            self._barrier()
        else:
            check(lib.qlb200_memcpy_h2d(self.ctx.h, C.c_void_p(self.in_ptr), host_in.ctypes.data, host_in.nbytes), "h2d")
        n_launch = (apply_fn or self.apply)()
        # after the apply's barrier EVERY GPU holds the whole result, so the download need not follow ownership (the edge ranks
        # own many small-sector rows: up to 1.5 x the mean output): rank r takes the r-th equal slice of the full buffer
        lo, hi = self.download_slice(host_out.size)
        if hi > lo:
            check(lib.qlb200_memcpy_d2h(self.ctx.h, host_out.ctypes.data + lo * es, C.c_void_p(self.full_ptr + lo * es), (hi - lo) * es), "d2h")
        self.ctx.sync()
        return n_launch

    def download_slice(self, n_elems: int):
        """Element range [lo, hi) of the full result this rank downloads in apply_host (equal slices, 16-byte aligned cuts)."""
        unit = max(1, 16 // self.dtype.itemsize)
        per = -(-n_elems // self.world)
        per = -(-per // unit) * unit
        return min(n_elems, self.rank * per), min(n_elems, (self.rank + 1) * per)

    def download_full(self, host: np.ndarray):
        check(lib.qlb200_memcpy_d2h(self.ctx.h, host.ctypes.data, C.c_void_p(self.full_ptr), host.nbytes), "qlb200_memcpy_d2h")

    def close(self):
        for pp in self.opened:
            lib.qlb200_ipc_close(self.ctx.h, C.c_void_p(pp))
        self.opened = []
        self.chain.close()
        if self.full_buf is not None:
            self.full_buf.free()
            self.full_buf = None
        if self.in_buf is not None:
            self.in_buf.free()
            self.in_buf = None
        if self.comm is not None:
            self.comm.close()
            self.comm = None
        if self.cplan:
            lib.qlb200_tplan_destroy(self.cplan)
            self.cplan = None
