"""Matrix-free axis operations: mirror of the reference's qlten::dmrg rank-2 axis application
(tensor_manipulation/dmrg/axis_ops.h) over the C ABI (qlb200_axis_*).

  apply_rank2_to_axis_preserve_order(x, op, axis)                  dmrg::ApplyRank2ToAxisPreserveOrder        (:2889-2992)
  apply_two_rank2_to_axes_preserve_order(x, op1, ax1, op2, ax2)    dmrg::ApplyTwoRank2ToAxesPreserveOrder     (:2994-3125)

`op` is a rank-2 tensor in the reference's {input_index, output_index} layout with op.indexes[0] == x.indexes[axis].inverse();
the result keeps x's axis order with the target axes replaced by the operators' output indexes.  Bosonic only, like the
reference.  Site-operator-sized blocks (edges <= 8) run in ONE kernel launch over the whole tensor; anything larger takes the
contraction path (contract, then move the new index back into place) -- same result, two launches plus a permute pass."""
import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import lib, check
from .contract import Context, contract, default_context, transpose
from .tensor import BlockSparseTensor, _dtype_code


class AxisPlan:
    """Block pairing, output topology and device tables of one axis application (plan once, execute many times)."""

    def __init__(self, ctx: Optional[Context], x: BlockSparseTensor, op1: BlockSparseTensor, axis1: int, op2: BlockSparseTensor = None,
                 axis2: int = -1, host_only: bool = False):
        ops = [(op1, axis1)] + ([(op2, axis2)] if op2 is not None else [])
        if x.rank and x.kind.fermionic:
            raise TypeError("axis operations are bosonic-only")
        for op, ax in ops:
            if not (0 <= ax < x.rank):
                raise ValueError("target axis out of range")
            if op.rank != 2:
                raise ValueError("rank2_op must have rank 2")
            if op.indexes[0] != x.indexes[ax].inverse():
                raise ValueError("rank2_op input index must be the inverse of the tensor axis")
        if op2 is not None and axis1 == axis2:
            raise ValueError("axes must be distinct")
        self.ctx, self.x, self.dtype = ctx, x, x.dtype
        self.sx, self.s1, self.s2 = x.shell(), op1.shell(), (op2.shell() if op2 is not None else None)
        h = C.c_void_p()
        check(lib.qlb200_axis_create(self.sx.ptr(), len(ops), self.s1.ptr(), axis1, self.s2.ptr() if self.s2 is not None else None,
                                     axis2 if op2 is not None else -1, C.byref(h)), "qlb200_axis_create")
        self.h = h
        self.out_indexes = list(x.indexes)
        for op, ax in ops:
            self.out_indexes[ax] = op.indexes[1]
        self.nterm = int(lib.qlb200_axis_nterm(h))
        self.plan = C.c_void_p()
        if not host_only:
            check(lib.qlb200_axis_plan_create(ctx.h, h, _dtype_code(self.dtype), C.byref(self.plan)), "qlb200_axis_plan_create")

    def result_shell(self) -> BlockSparseTensor:
        out = BlockSparseTensor(self.out_indexes, self.dtype)
        n = int(lib.qlb200_axis_out_nblk(self.h))
        if n:
            coors = np.zeros((n, out.rank), np.uint32)
            check(lib.qlb200_axis_out_blocks(self.h, None, coors.ctypes.data_as(C.POINTER(C.c_uint32)), None, None), "qlb200_axis_out_blocks")
            out.set_blocks(coors)
            assert out.data.size == int(lib.qlb200_axis_out_elems(self.h))
        return out

    def bytes(self):
        r, w = C.c_uint64(), C.c_uint64()
        check(lib.qlb200_axis_plan_bytes(self.plan, C.byref(r), C.byref(w)), "qlb200_axis_plan_bytes")
        return int(r.value), int(w.value)

    def execute_host(self, x, op1, op2, out):
        check(lib.qlb200_axis_execute(self.ctx.h, self.plan, x.ctypes.data, op1.ctypes.data, op2.ctypes.data if op2 is not None else None,
                                      out.ctypes.data, _lib.MEM_HOST), "qlb200_axis_execute")

    def execute_device(self, x_ptr, op1_ptr, op2_ptr, out_ptr):
        check(lib.qlb200_axis_execute(self.ctx.h, self.plan, C.c_void_p(x_ptr), C.c_void_p(op1_ptr), C.c_void_p(op2_ptr) if op2_ptr else None,
                                      C.c_void_p(out_ptr), _lib.MEM_DEVICE), "qlb200_axis_execute")

    def close(self):
        if self.plan:
            lib.qlb200_axis_plan_destroy(self.plan)
            self.plan = None
        if self.h:
            lib.qlb200_axis_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _via_contraction(x, op, axis, ctx):
    """out = Contract(x, op, {{axis}, {0}}) with the new (last) index moved back to position `axis`."""
    c = contract(x, op, ([axis], [0]), ctx)
    r = x.rank
    order = list(range(axis)) + [r - 1] + list(range(axis, r - 1))
    return c if order == list(range(r)) else transpose(c, order, ctx)


def apply_rank2_to_axis_preserve_order(x: BlockSparseTensor, rank2_op: BlockSparseTensor, target_axis: int, ctx: Context = None) -> BlockSparseTensor:
    ctx = ctx or default_context()
    try:
        plan = AxisPlan(ctx, x, rank2_op, target_axis)
    except _lib.QLB200Error as e:
        if "contraction path" not in str(e):
            raise
        return _via_contraction(x, rank2_op, target_axis, ctx)
    try:
        out = plan.result_shell()
        if out.data.size:
            plan.execute_host(x.data, rank2_op.data, None, out.data)
        return out
    finally:
        plan.close()


def apply_two_rank2_to_axes_preserve_order(x: BlockSparseTensor, op1: BlockSparseTensor, axis1: int, op2: BlockSparseTensor, axis2: int,
                                           ctx: Context = None) -> BlockSparseTensor:
    ctx = ctx or default_context()
    try:
        plan = AxisPlan(ctx, x, op1, axis1, op2, axis2)
    except _lib.QLB200Error as e:
        if "contraction path" not in str(e):
            raise
        return _via_contraction(_via_contraction(x, op1, axis1, ctx), op2, axis2, ctx)
    try:
        out = plan.result_shell()
        if out.data.size:
            plan.execute_host(x.data, op1.data, op2.data, out.data)
        return out
    finally:
        plan.close()
