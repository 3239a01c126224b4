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
        """End-to-end apply: input H2D, all steps, result D2H, synchronise."""
        self.buf[in_name].upload(host_in)
        self.apply_device()
        self.buf[out_name].download(host_out)
        self.ctx.sync()

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
        for p in self.plans:
            p.close()
        for m in self.matches:
            m.close()
        for b in self.buf.values():
            b.free()
        self.plans, self.matches, self.buf = [], [], {}


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
    exchange=None: no exchange at all (time one rank's share on a single GPU)."""

    def __init__(self, ctx: Context, tensors: Dict[str, BlockSparseTensor], steps, name: str, axis: int, dtype,
                 world: int, rank: int, group=None, flags: int = _lib.PLAN_DETERMINISTIC, exchange="auto", peers=None):
        import torch
        from .sharding import shard_chain
        self.torch, self.group = torch, group
        self.ctx, self.world, self.rank = ctx, world, rank
        mine, self.info = shard_chain(tensors, steps, name, axis, world, rank, dtype)
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
        self.chain = ContractionChain(ctx, mine, steps, dtype, flags, external={self.out_name: self.local.data_ptr()},
                                      last_flags=stagger)
        full_bytes = max(self.info.full_elems, 1) * self.dtype.itemsize
        self.symm = None
        if exchange in ("auto", "multicast") and world > 1 and peers is None:
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else torch.distributed.group.WORLD
            with torch.cuda.device(dev):
                self.symm_buf = symm_mem.empty(max(self.info.full_elems, 1), dtype=tdt, device=dev)
                self.symm = symm_mem.rendezvous(self.symm_buf, grp)
            if not int(self.symm.multicast_ptr):
                if exchange == "multicast":
                    raise RuntimeError("exchange='multicast': this NVLink domain offers no multicast (NVLS) mapping")
                self.symm, self.symm_buf, exchange = None, None, "fused"
            else:
                exchange = "multicast"
                self.full_ptr = int(self.symm_buf.data_ptr())
                self.mc_ptr = int(self.symm.multicast_ptr)
                my = self.info.slabs[rank]
                self.chain.plans[-1].remap_output([s.local_offset for s in my], [s.full_offset for s in my])
        elif exchange == "auto":
            exchange = "fused"
        self.exchange = exchange
        if exchange == "multicast":
            pass
        elif exchange == "fused":
            # the full result lives in a cudaMalloc'ed buffer of its own so that it can be exported over CUDA IPC
            self.full_buf = DeviceBuffer(ctx, full_bytes)
            self.full_ptr = self.full_buf.ptr
            my = self.info.slabs[rank]
            self.chain.plans[-1].remap_output([s.local_offset for s in my], [s.full_offset for s in my])
            if peers is not None:                       # caller-supplied replicas (single-process tests)
                self.peer_ptrs = [self.full_ptr] + [int(x) for x in peers]
            elif world > 1:
                handle = C.create_string_buffer(64)
                check(lib.qlb200_ipc_export(ctx.h, C.c_void_p(self.full_ptr), handle), "qlb200_ipc_export")
                handles = [None] * world
                torch.distributed.all_gather_object(handles, bytes(handle.raw), group=group)
                self.peer_ptrs = [self.full_ptr]
                for r in range(world):
                    if r == rank:
                        continue
                    pp = C.c_void_p()
                    check(lib.qlb200_ipc_open(ctx.h, handles[r], C.byref(pp)), "qlb200_ipc_open")
                    self.opened.append(pp.value)
                    self.peer_ptrs.append(pp.value)
            else:
                self.peer_ptrs = [self.full_ptr]
            self.flag = torch.zeros(1, dtype=torch.float32, device=dev)
        else:
            self.gathered = torch.zeros(world * self.stride, dtype=tdt, device=dev)
            self.full = torch.zeros(max(self.info.full_elems, 1), dtype=tdt, device=dev)
            self.full_ptr = self.full.data_ptr()
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
        if self.exchange == "multicast":
            n = 0
            for (lhs, rhs, _, out), plan in zip(ch.steps[:-1], ch.plans[:-1]):
                plan.execute_device(ch.buf[lhs].ptr, ch.buf[rhs].ptr, ch.buf[out].ptr)
                n += self.ctx.launch_count()
            lhs, rhs, _, _ = ch.steps[-1]
            ch.plans[-1].execute_mcast(ch.buf[lhs].ptr, ch.buf[rhs].ptr, self.mc_ptr)
            n += self.ctx.launch_count()
            mark("compute")
            self.symm.barrier()      # every rank's tiles have landed in every replica
            mark("exchange")
            return n
        if self.exchange == "fused":
            n = 0
            for (lhs, rhs, _, out), plan in zip(ch.steps[:-1], ch.plans[:-1]):
                plan.execute_device(ch.buf[lhs].ptr, ch.buf[rhs].ptr, ch.buf[out].ptr)
                n += self.ctx.launch_count()
            lhs, rhs, _, _ = ch.steps[-1]
            ch.plans[-1].execute_bcast(ch.buf[lhs].ptr, ch.buf[rhs].ptr, self.peer_ptrs)
            n += self.ctx.launch_count()
            mark("compute")
            if self.world > 1 and self.opened:
                self.torch.distributed.all_reduce(self.flag, group=self.group)    # barrier: every peer's tiles have landed
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
        check(lib.qlb200_copy_execute(self.ctx.h, self.cplan, C.c_void_p(self.gathered.data_ptr()), C.c_void_p(self.full_ptr)),
              "qlb200_copy_execute")
        mark("exchange")
        return n + 1

    def capture(self) -> CapturedGraph:
        """Local steps + exchange + barrier as one CUDA graph (the torch stream must be the context's stream)."""
        return CapturedGraph(self.ctx, self.apply)

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
        if self.cplan:
            lib.qlb200_tplan_destroy(self.cplan)
            self.cplan = None
