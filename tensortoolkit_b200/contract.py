"""Python face of the contraction path; every call goes through the C ABI (include/qlb200.h).

  contract(a, b, axes)            mirrors qlten::Contract (tensor_manipulation/ten_ctrct.h:277-290)
  contract_1sector(a, ax, s, b, axes)  mirrors qlten::dmrg::Contract1Sector (dmrg/contract_1sector.h:211-228)
  transpose(t, order)             mirrors QLTensor::Transpose (qltensor_impl.h:449-464)

`ContractionPlan` exposes the plan/execute split for device-resident, repeated use (Lanczos-style
loops, benchmarks): match once, build the descriptor tables once, execute many times.
"""
import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import lib, check
from .tensor import BlockSparseTensor, _dtype_code


class Context:
    """One qlb200_ctx (one device, one stream). Fails loudly without a B200."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        check(lib.qlb200_ctx_create(device, C.byref(h)), "qlb200_ctx_create")
        self.h = h
        self.device = device

    def sync(self):
        check(lib.qlb200_ctx_sync(self.h), "qlb200_ctx_sync")

    def stream(self) -> int:
        return lib.qlb200_ctx_stream(self.h) or 0

    def set_stream(self, cuda_stream: int):
        check(lib.qlb200_ctx_set_stream(self.h, C.c_void_p(cuda_stream)), "qlb200_ctx_set_stream")

    def launch_count(self) -> int:
        return int(lib.qlb200_ctx_launch_count(self.h))

    def close(self):
        if self.h:
            lib.qlb200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _i32(v):
    return (C.c_int32 * len(v))(*[int(x) for x in v])


class Match:
    """Result of the host sector matcher (qlb200_match)."""

    def __init__(self, a: BlockSparseTensor, b: BlockSparseTensor, axes, one_sector=None, contiguous=None):
        """axes = (a axes, b axes) for the general contraction; contiguous = (a_start, b_start, size) selects the
        block pairing / result layout of qlten::ContractContiguousAxes instead (axes is then ignored)."""
        if contiguous is not None:
            a_start, b_start, size = (int(x) for x in contiguous)
            if not (0 <= a_start < a.rank and 0 <= b_start < b.rank and 0 <= size <= min(a.rank, b.rank)):
                raise ValueError("bad contiguous axis range")
            axes = ([(a_start + i) % a.rank for i in range(size)], [(b_start + i) % b.rank for i in range(size)])
        axes_a, axes_b = list(axes[0]), list(axes[1])
        if len(axes_a) != len(axes_b):
            raise ValueError("axes_set must pair axes of A with axes of B")
        for x, y in zip(axes_a, axes_b):
            if not (0 <= x < a.rank and 0 <= y < b.rank) or a.indexes[x] != b.indexes[y].inverse():
                raise ValueError(f"contracted indexes A[{x}] and B[{y}] do not match (need A[x] == InverseIndex(B[y]))")
        self.a, self.b = a, b
        self.axes = (axes_a, axes_b)
        self.sa, self.sb = a.shell(), b.shell()
        h = C.c_void_p()
        if contiguous is not None:
            rc = lib.qlb200_match_create_contiguous(self.sa.ptr(), self.sb.ptr(), a_start, b_start, size, C.byref(h))
        elif one_sector is None:
            rc = lib.qlb200_match_create(self.sa.ptr(), self.sb.ptr(), len(axes_a), _i32(axes_a), _i32(axes_b), C.byref(h))
        else:
            rc = lib.qlb200_match_create_1sector(self.sa.ptr(), int(one_sector[0]), int(one_sector[1]), self.sb.ptr(),
                                                 len(axes_a), _i32(axes_a), _i32(axes_b), C.byref(h))
        check(rc, "qlb200_match_create")
        self.h = h
        self.c_rank = lib.qlb200_match_c_rank(h)
        self.c_nblk = int(lib.qlb200_match_c_nblk(h))
        self.c_elems = int(lib.qlb200_match_c_elems(h))
        self.ntask = int(lib.qlb200_match_ntask(h))
        self.is_scalar = bool(lib.qlb200_match_is_scalar(h))
        buf = (C.c_int32 * _lib.QLB200_MAX_RANK)()
        saved_a = [int(buf[i]) for i in range(lib.qlb200_match_saved_axes(h, 0, buf))]
        saved_b = [int(buf[i]) for i in range(lib.qlb200_match_saved_axes(h, 1, buf))]
        self.saved_axes = (saved_a, saved_b)
        self.c_indexes = [a.indexes[i] for i in saved_a] + [b.indexes[i] for i in saved_b]

    def perm(self, which: int):
        rank = self.a.rank if which == 0 else self.b.rank
        out = (C.c_int32 * rank)()
        need = lib.qlb200_match_perm(self.h, which, out)
        return list(out), bool(need)

    def c_blocks(self):
        n, r = self.c_nblk, self.c_rank
        idx = np.zeros(n, np.uint64); off = np.zeros(n, np.uint64)
        coors = np.zeros((n, r), np.uint32); shape = np.zeros((n, r), np.uint32)
        check(lib.qlb200_match_c_blocks(self.h, idx.ctypes.data_as(C.POINTER(C.c_uint64)),
                                        coors.ctypes.data_as(C.POINTER(C.c_uint32)),
                                        shape.ctypes.data_as(C.POINTER(C.c_uint32)),
                                        off.ctypes.data_as(C.POINTER(C.c_uint64))), "qlb200_match_c_blocks")
        return idx, coors, shape, off

    def tasks(self, sorted_by_c: bool = False):
        arr = (_lib.Task * self.ntask)()
        check(lib.qlb200_match_tasks(self.h, 1 if sorted_by_c else 0, arr), "qlb200_match_tasks")
        return arr

    def cost(self, dtype) -> _lib.Cost:
        c = _lib.Cost()
        check(lib.qlb200_estimate_cost(self.h, _dtype_code(dtype), C.byref(c)), "qlb200_estimate_cost")
        return c

    def result_shell(self, dtype) -> BlockSparseTensor:
        c = BlockSparseTensor(self.c_indexes, dtype)
        if not self.is_scalar and self.c_nblk:
            _, coors, _, _ = self.c_blocks()
            c.set_blocks(coors)
        elif self.is_scalar:
            c.data = np.zeros(1 if self.ntask else 0, dtype=c.dtype)
        return c

    def close(self):
        if self.h:
            lib.qlb200_match_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ContractionPlan:
    """Device descriptor tables of one contraction (qlb200_plan)."""

    def __init__(self, ctx: Context, match: Match, dtype, flags: int = _lib.PLAN_DETERMINISTIC):
        self.ctx, self.match = ctx, match
        self.dtype = np.dtype(dtype)
        h = C.c_void_p()
        check(lib.qlb200_plan_create(ctx.h if ctx is not None else None, match.h, match.sa.ptr(), match.sb.ptr(), _dtype_code(dtype), flags, C.byref(h)),
              "qlb200_plan_create")
        self.h = h

    def stats(self) -> _lib.PlanStats:
        s = _lib.PlanStats()
        check(lib.qlb200_plan_get_stats(self.h, C.byref(s)), "qlb200_plan_get_stats")
        return s

    def partition(self, world: int, rank: int):
        check(lib.qlb200_plan_partition(self.h, world, rank), "qlb200_plan_partition")

    def c_ranges(self):
        n = int(lib.qlb200_plan_c_range_count(self.h))
        off = np.zeros(n, np.uint64); ln = np.zeros(n, np.uint64)
        check(lib.qlb200_plan_c_ranges(self.h, off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                       ln.ctypes.data_as(C.POINTER(C.c_uint64))), "qlb200_plan_c_ranges")
        return off, ln

    def execute_host(self, a: np.ndarray, b: np.ndarray, c: np.ndarray):
        check(lib.qlb200_execute(self.ctx.h, self.h, a.ctypes.data, b.ctypes.data, c.ctypes.data, _lib.MEM_HOST), "qlb200_execute")

    def units(self):
        """(list of work units in launch order, (tile rows, tile cols, k per stage)) -- qlb200_plan_units."""
        n = int(lib.qlb200_plan_units(self.h, 0, None, None, None, None))
        arr = (_lib.Unit * max(n, 1))()
        bm, bn, bk = C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib.qlb200_plan_units(self.h, n, arr, C.byref(bm), C.byref(bn), C.byref(bk))
        return [arr[i] for i in range(n)], (bm.value, bn.value, bk.value)

    def items(self):
        """Work items of the narrow-pair kernel: list of (group, row0, rows, n) -- qlb200_plan_items."""
        n = int(lib.qlb200_plan_items(self.h, 0, None))
        arr = (_lib.Item * max(n, 1))()
        lib.qlb200_plan_items(self.h, n, arr)
        return [(arr[i].group, arr[i].row0, arr[i].rows, arr[i].n) for i in range(n)]

    def segments(self):
        """Stream-K plans: unit range boundaries per CTA (qlb200_plan_segments); [] for dynamically scheduled plans."""
        n = int(lib.qlb200_plan_segments(self.h, 0, None))
        arr = (C.c_uint32 * max(n, 1))()
        lib.qlb200_plan_segments(self.h, n, arr)
        return [int(arr[i]) for i in range(n)]

    def execute_device(self, a_ptr: int, b_ptr: int, c_ptr: int):
        check(lib.qlb200_execute(self.ctx.h, self.h, C.c_void_p(a_ptr), C.c_void_p(b_ptr), C.c_void_p(c_ptr), _lib.MEM_DEVICE),
              "qlb200_execute")

    def execute_mcast(self, a_ptr: int, b_ptr: int, c_multicast_ptr: int):
        """Permute + GEMM with the output stored through an NVSwitch multicast mapping of the result (multimem.st)."""
        check(lib.qlb200_execute_mcast(self.ctx.h, self.h, C.c_void_p(a_ptr), C.c_void_p(b_ptr), C.c_void_p(c_multicast_ptr)),
              "qlb200_execute_mcast")

    def execute_bcast(self, a_ptr: int, b_ptr: int, c_ptrs):
        """Device-resident execute whose output tiles are stored to every buffer of `c_ptrs` (this GPU's and
        its NVLink peers') from inside the GEMM epilogue."""
        arr = (C.c_void_p * len(c_ptrs))(*[int(x) for x in c_ptrs])
        check(lib.qlb200_execute_bcast(self.ctx.h, self.h, C.c_void_p(a_ptr), C.c_void_p(b_ptr), arr, len(c_ptrs)), "qlb200_execute_bcast")

    def remap_output(self, from_off, to_off):
        f = np.ascontiguousarray(np.asarray(from_off, np.uint64)); t = np.ascontiguousarray(np.asarray(to_off, np.uint64))
        check(lib.qlb200_plan_remap_output(self.h, len(f), f.ctypes.data_as(C.POINTER(C.c_uint64)), t.ctypes.data_as(C.POINTER(C.c_uint64))),
              "qlb200_plan_remap_output")

    def operand_block(self, which: int, ord_: int):
        """None when the GEMM reads block `ord_` of operand `which` in place, else its element offset in the permuted workspace."""
        off = C.c_uint64()
        rc = lib.qlb200_plan_operand_block(self.h, which, ord_, C.byref(off))
        if rc < 0:
            check(rc, "qlb200_plan_operand_block")
        return int(off.value) if rc == 1 else None

    def read_workspace(self, which: int, elem_off: int, elems: int) -> np.ndarray:
        """Permuted copy of operand `which` (elements [elem_off, elem_off + elems) of its workspace region) after execute_permute."""
        out = np.empty(elems, self.dtype)
        check(lib.qlb200_plan_read_workspace(self.ctx.h, self.h, which, elem_off, elems, out.ctypes.data), "qlb200_plan_read_workspace")
        return out

    def execute_permute(self, a_ptr: int, b_ptr: int):
        check(lib.qlb200_execute_permute(self.ctx.h, self.h, C.c_void_p(a_ptr), C.c_void_p(b_ptr)), "qlb200_execute_permute")

    def execute_gemm(self, a_ptr: int, b_ptr: int, c_ptr: int):
        check(lib.qlb200_execute_gemm(self.ctx.h, self.h, C.c_void_p(a_ptr), C.c_void_p(b_ptr), C.c_void_p(c_ptr)),
              "qlb200_execute_gemm")

    def close(self):
        if self.h:
            lib.qlb200_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RawPlan(ContractionPlan):
    """Descriptor-table plan without QLTensor shells (qlb200_plan_create_raw): the caller lists block
    shapes / offsets of A and B, one permutation per operand and the task table.  `tasks` is a list of
    dicts with a_ord, b_ord, a_off, b_off, c_off, m, k, n, sign, first."""

    def __init__(self, ctx: Context, dtype, a_rank, a_perm, a_shape, a_off, b_rank, b_perm, b_shape, b_off, tasks, c_elems,
                 flags: int = _lib.PLAN_DETERMINISTIC):
        self.ctx, self.match = ctx, None
        self.dtype = np.dtype(dtype)
        ash = np.ascontiguousarray(np.asarray(a_shape, np.uint32).reshape(-1))
        bsh = np.ascontiguousarray(np.asarray(b_shape, np.uint32).reshape(-1))
        aof = np.ascontiguousarray(np.asarray(a_off, np.uint64)); bof = np.ascontiguousarray(np.asarray(b_off, np.uint64))
        if isinstance(tasks, np.ndarray):       # structured array with the layout of qlb200_task
            assert tasks.dtype.itemsize == C.sizeof(_lib.Task)
            tarr, nt = tasks.ctypes.data_as(C.POINTER(_lib.Task)), len(tasks)
        else:
            tarr = (_lib.Task * max(len(tasks), 1))()
            for i, d in enumerate(tasks):
                t = tarr[i]
                t.a_ord, t.b_ord, t.c_ord = d["a_ord"], d["b_ord"], d.get("c_ord", 0)
                t.a_off, t.b_off, t.c_off = d["a_off"], d["b_off"], d["c_off"]
                t.m, t.k, t.n, t.sign, t.first = d["m"], d["k"], d["n"], d.get("sign", 1), d.get("first", 0)
            nt = len(tasks)
        h = C.c_void_p()
        u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        check(lib.qlb200_plan_create_raw(
            ctx.h if ctx is not None else None, _dtype_code(dtype), flags, a_rank, (C.c_int32 * a_rank)(*a_perm), len(aof), ash.ctypes.data_as(u32p),
            aof.ctypes.data_as(u64p), b_rank, (C.c_int32 * b_rank)(*b_perm), len(bof), bsh.ctypes.data_as(u32p),
            bof.ctypes.data_as(u64p), nt, tarr, int(c_elems), C.byref(h)), "qlb200_plan_create_raw")
        self.h = h


def _run(a, b, match, ctx):
    if a.dtype != b.dtype:
        # mixed real/complex promotes like the reference (ten_ctrct.h:292-350)
        if a.dtype != np.complex128:
            a2 = BlockSparseTensor(a.indexes, np.complex128); a2.set_blocks(a.blk_coors, a.data.astype(np.complex128)); a = a2
        if b.dtype != np.complex128:
            b2 = BlockSparseTensor(b.indexes, np.complex128); b2.set_blocks(b.blk_coors, b.data.astype(np.complex128)); b = b2
    c = match.result_shell(a.dtype)
    if match.ntask == 0:
        return c
    ctx = ctx or default_context()
    plan = ContractionPlan(ctx, match, a.dtype)
    try:
        plan.execute_host(a.data, b.data, c.data)
    finally:
        plan.close()
    return c


def contract(a: BlockSparseTensor, b: BlockSparseTensor, axes: Sequence[Sequence[int]], ctx: Context = None) -> BlockSparseTensor:
    """C = Contract(A, B, {{a axes}, {b axes}}): C's indexes are A's free indexes then B's free indexes."""
    m = Match(a, b, axes)
    try:
        return _run(a, b, m, ctx)
    finally:
        m.close()


def contract_contiguous_axes(a: BlockSparseTensor, b: BlockSparseTensor, a_ctrct_axes_start: int, b_ctrct_axes_start: int,
                             ctrct_axes_size: int, ctx: Context = None) -> BlockSparseTensor:
    """qlten::ContractContiguousAxes (tensor_manipulation/contract_contiguous_axes.h:849-873): contract axes
    (a_start + i) % rank_a of A with (b_start + i) % rank_b of B; the result's indexes are A's free indexes in cyclic
    order starting behind the contracted range, then B's likewise.  Every block is read in place by the grouped GEMM
    (a cyclic rotation is a 2-D transposition of the block), so no transpose pass runs."""
    m = Match(a, b, None, contiguous=(a_ctrct_axes_start, b_ctrct_axes_start, ctrct_axes_size))
    try:
        return _run(a, b, m, ctx)
    finally:
        m.close()


class AccumulateLayoutMismatch(ValueError):
    """The existing output cannot take the contraction result (the reference's detail::ContractAccumulateLayoutMismatch,
    contract_contiguous_axes.h:83-88): incompatible indexes or block shapes, or missing blocks when expansion is not allowed."""


def contract_tail_head_contiguous_accumulate(a: BlockSparseTensor, b: BlockSparseTensor, a_ctrct_axes_start: int,
                                             b_ctrct_axes_start: int, ctrct_axes_size: int, alpha, beta,
                                             c: Optional[BlockSparseTensor] = None, ctx: Context = None, stats: dict = None,
                                             allow_output_topology_expansion: bool = True) -> BlockSparseTensor:
    """qlten::ContractTailHeadContiguousAccumulate (tensor_manipulation/contract_contiguous_axes.h:954-1000):
        c  <-  beta * c + alpha * ContractContiguousAxes(a, b, ...)
    without a temporary result tensor.  `c` None = a default tensor (beta must be 0).  An existing `c` may hold more blocks
    than the contraction produces (scaled by beta only) or fewer (rebuilt on the union topology).  Returns the new c
    (the argument is not modified); `stats`, if given, receives the reference's ContiguousContractStats counters."""
    if c is a or c is b:
        raise ValueError("ContractTailHeadContiguousAccumulate does not support aliasing between output and input tensors")
    m = Match(a, b, None, contiguous=(a_ctrct_axes_start, b_ctrct_axes_start, ctrct_axes_size))
    acc = C.c_void_p()
    plan = C.c_void_p()
    try:
        dtype = a.dtype
        if c is not None and list(c.indexes) != list(m.c_indexes):
            raise AccumulateLayoutMismatch("output indexes are not compatible with the contraction result")
        cplx = np.dtype(dtype) == np.complex128
        if not cplx and (np.iscomplexobj(alpha) and complex(alpha).imag != 0 or np.iscomplexobj(beta) and complex(beta).imag != 0):
            raise TypeError("complex alpha / beta on real tensors")
        al = (C.c_double * 2)(complex(alpha).real, complex(alpha).imag)
        be = (C.c_double * 2)(complex(beta).real, complex(beta).imag)
        sh = c.shell() if c is not None else None
        has_data = int(c is not None and c.data.size > 0)
        rc = lib.qlb200_accum_create(m.h, sh.ptr() if sh is not None else None, has_data, int(allow_output_topology_expansion),
                                     _dtype_code(dtype), al, be, C.byref(acc))
        if rc == _lib.ERR_LAYOUT:
            raise AccumulateLayoutMismatch(lib.qlb200_last_error().decode())
        if rc == _lib.ERR_ARG:
            raise ValueError(lib.qlb200_last_error().decode())
        check(rc, "qlb200_accum_create")
        if stats is not None:
            st = _lib.AccumStats()
            check(lib.qlb200_accum_get_stats(acc, C.byref(st)), "qlb200_accum_get_stats")
            stats.update(st.as_dict())
        out = BlockSparseTensor(m.c_indexes, dtype)
        n = int(lib.qlb200_accum_nblk(acc))
        if m.is_scalar:
            out.data = np.zeros(1, dtype=out.dtype)
        elif n:
            coors = np.zeros((n, m.c_rank), np.uint32)
            check(lib.qlb200_accum_blocks(acc, None, coors.ctypes.data_as(C.POINTER(C.c_uint32)), None, None, None, None), "qlb200_accum_blocks")
            out.set_blocks(coors)
            assert out.data.size == int(lib.qlb200_accum_elems(acc))
        if out.data.size == 0:
            return out
        ctx = ctx or default_context()
        check(lib.qlb200_plan_create_accum(ctx.h, m.h, acc, _dtype_code(dtype), _lib.PLAN_DETERMINISTIC, C.byref(plan)), "qlb200_plan_create_accum")
        old = c.data.ctypes.data if (c is not None and c.data.size) else None
        check(lib.qlb200_execute_accum(ctx.h, plan, a.data.ctypes.data, b.data.ctypes.data, old, out.data.ctypes.data, _lib.MEM_HOST),
              "qlb200_execute_accum")
        return out
    finally:
        if plan:
            lib.qlb200_plan_destroy(plan)
        if acc:
            lib.qlb200_accum_destroy(acc)
        m.close()


def try_contract_tail_head_contiguous_accumulate(a, b, a_ctrct_axes_start, b_ctrct_axes_start, ctrct_axes_size, alpha, beta,
                                                 c=None, ctx: Context = None, stats: dict = None):
    """qlten::TryContractTailHeadContiguousAccumulate (contract_contiguous_axes.h:1002-1041): the no-expansion probe.
    Returns (True, new c) or (False, c unchanged) on a layout mismatch; other errors still raise."""
    try:
        return True, contract_tail_head_contiguous_accumulate(a, b, a_ctrct_axes_start, b_ctrct_axes_start, ctrct_axes_size, alpha, beta,
                                                              c, ctx, stats, allow_output_topology_expansion=False)
    except AccumulateLayoutMismatch:
        if stats is not None:
            stats.clear()
        return False, c


def contract_1sector(a, idx_a: int, qn_sector_idx_a: int, b, axes, ctx: Context = None) -> BlockSparseTensor:
    if idx_a in list(axes[0]):
        raise ValueError("the split index must be a free index of A")
    m = Match(a, b, axes, one_sector=(idx_a, qn_sector_idx_a))
    try:
        return _run(a, b, m, ctx)
    finally:
        m.close()


def transpose(t: BlockSparseTensor, order: Sequence[int], ctx: Context = None) -> BlockSparseTensor:
    """Whole-tensor index permutation; fermionic tensors pick up the reorder sign per block."""
    order = [int(x) for x in order]
    if sorted(order) != list(range(t.rank)):
        raise ValueError("order is not a permutation")
    ctx = ctx or default_context()
    sh = t.shell()
    h = C.c_void_p()
    check(lib.qlb200_tplan_create(ctx.h, sh.ptr(), _i32(order), _dtype_code(t.dtype), C.byref(h)), "qlb200_tplan_create")
    try:
        out = BlockSparseTensor([t.indexes[i] for i in order], t.dtype)
        n = int(lib.qlb200_tplan_nblk(h))
        coors = np.zeros((n, t.rank), np.uint32)
        check(lib.qlb200_tplan_blocks(h, None, coors.ctypes.data_as(C.POINTER(C.c_uint32)), None, None, None), "qlb200_tplan_blocks")
        out.set_blocks(coors)
        if n:
            check(lib.qlb200_transpose_execute(ctx.h, h, t.data.ctypes.data, out.data.ctypes.data, _lib.MEM_HOST),
                  "qlb200_transpose_execute")
        return out
    finally:
        lib.qlb200_tplan_destroy(h)
