// qlten_b200/axis_ops.h -- drop-in adapter for the matrix-free axis operations of TensorToolkit's DMRG tool box:
//
//   qlten::b200::dmrg::ApplyRank2ToAxisPreserveOrder(input, rank2_op, target_axis, out)          same signature / semantics as
//       qlten::dmrg::ApplyRank2ToAxisPreserveOrder    (include/qlten/tensor_manipulation/dmrg/axis_ops.h:2889-2992)
//   qlten::b200::dmrg::ApplyTwoRank2ToAxesPreserveOrder(input, op1, axis1, op2, axis2, out)      as
//       qlten::dmrg::ApplyTwoRank2ToAxesPreserveOrder (:2994-3125)
//
// Both go through the qlb200_axis_* entry points of include/qlb200.h: one kernel launch over the whole tensor, no single-axis
// intermediate.  Operator blocks wider than 8 (not site operators) take the contraction path: qlten::b200::Contract followed
// by qlten::b200::Transpose of the new index into place.  Only public reference API is used; bosonic quantum numbers only
// (static_assert, like the reference).  The stats out-parameters of the reference count its own GEMM calls and are not filled.
#ifndef QLTEN_B200_AXIS_OPS_H
#define QLTEN_B200_AXIS_OPS_H

#include "qlten_b200/contract.h"
#include "qlten/tensor_manipulation/dmrg/axis_ops.h"

namespace qlten {
namespace b200 {
namespace dmrg {

namespace detail {

struct AxisGuard {
  qlb200_axis *a = nullptr;
  qlb200_axis_plan *p = nullptr;
  ~AxisGuard() { if (p) qlb200_axis_plan_destroy(p); if (a) qlb200_axis_destroy(a); }
};

template<typename ElemT, typename QNT>
void CheckOp(const QLTensor<ElemT, QNT> &input, const QLTensor<ElemT, QNT> &op, size_t axis, const char *func) {
  if (input.IsDefault() || axis >= input.Rank()) throw std::invalid_argument(std::string(func) + ": invalid input / axis.");
  if (op.IsDefault()) throw std::invalid_argument(std::string(func) + ": rank2_op must not be default.");
  if (op.Rank() != 2) throw std::invalid_argument(std::string(func) + ": rank2_op must have rank 2.");
  if (op.GetIndex(0) != InverseIndex(input.GetIndex(axis)))
    throw std::invalid_argument(std::string(func) + ": rank2_op input index must be the inverse of the tensor axis.");
}

/// contraction path for operator blocks too large for the axis kernel: same result, more passes
template<typename ElemT, typename QNT>
void ViaContraction(const QLTensor<ElemT, QNT> &input, const QLTensor<ElemT, QNT> &op, size_t axis, QLTensor<ElemT, QNT> &out, qlb200_ctx *ctx) {
  QLTensor<ElemT, QNT> c;
  qlten::b200::Contract(&input, &op, {{axis}, {0}}, &c, ctx);
  const size_t r = input.Rank();
  std::vector<size_t> order;
  for (size_t i = 0; i < axis; ++i) order.push_back(i);
  order.push_back(r - 1);
  for (size_t i = axis; i + 1 < r; ++i) order.push_back(i);
  if (!c.IsDefault() && !c.IsScalar()) qlten::b200::Transpose(&c, order, ctx);
  out = std::move(c);
}

template<typename ElemT, typename QNT>
bool RunAxis(const QLTensor<ElemT, QNT> &input, const QLTensor<ElemT, QNT> *op1, size_t axis1, const QLTensor<ElemT, QNT> *op2, size_t axis2,
             const IndexVec<QNT> &out_indexes, QLTensor<ElemT, QNT> &out, qlb200_ctx *ctx) {
  using qlten::b200::detail::Check;
  qlten::b200::detail::ShellHolder si, s1, s2;
  qlten::b200::detail::FillShell(input, si);
  qlten::b200::detail::FillShell(*op1, s1);
  if (op2 != nullptr) qlten::b200::detail::FillShell(*op2, s2);
  AxisGuard g;
  Check(qlb200_axis_create(&si.shell, op2 != nullptr ? 2 : 1, &s1.shell, static_cast<int32_t>(axis1), op2 != nullptr ? &s2.shell : nullptr,
                           static_cast<int32_t>(axis2), &g.a), "axis_create");
  const int rc = qlb200_axis_plan_create(ctx, g.a, qlten::b200::detail::DTypeOf<ElemT>::value, &g.p);
  if (rc == QLB200_ERR_UNSUPPORTED) return false;
  Check(rc, "axis_plan_create");
  out = QLTensor<ElemT, QNT>(out_indexes);
  const uint64_t nblk = qlb200_axis_out_nblk(g.a);
  if (nblk == 0) return true;
  const size_t rank = input.Rank();
  std::vector<uint64_t> blk_idx(nblk);
  std::vector<uint32_t> coors(nblk * rank);
  Check(qlb200_axis_out_blocks(g.a, blk_idx.data(), coors.data(), nullptr, nullptr), "axis_out_blocks");
  std::vector<size_t> idxs(blk_idx.begin(), blk_idx.end());
  std::vector<CoorsT> coors_s(nblk, CoorsT(rank));
  for (uint64_t b = 0; b < nblk; ++b)
    for (size_t i = 0; i < rank; ++i) coors_s[b][i] = coors[b * rank + i];
  out.GetBlkSparDataTen().DataBlksInsert(idxs, coors_s, true);
  Check(qlb200_axis_execute(ctx, g.p, input.GetBlkSparDataTen().GetActualRawDataPtr(), op1->GetBlkSparDataTen().GetActualRawDataPtr(),
                            op2 != nullptr ? op2->GetBlkSparDataTen().GetActualRawDataPtr() : nullptr,
                            out.GetBlkSparDataTen().GetActualRawDataPtr(), QLB200_MEM_HOST), "axis_execute");
  return true;
}

}  // namespace detail

template<typename ElemT, typename QNT>
void ApplyRank2ToAxisPreserveOrder(const QLTensor<ElemT, QNT> &input, const QLTensor<ElemT, QNT> &rank2_op, size_t target_axis,
                                   QLTensor<ElemT, QNT> &out, qlb200_ctx *ctx = nullptr) {
  static_assert(!Fermionicable<QNT>::IsFermionic(), "ApplyRank2ToAxisPreserveOrder is bosonic-only.");
  if (&input == &out || &rank2_op == &out) throw std::invalid_argument("ApplyRank2ToAxisPreserveOrder: output aliasing is not allowed.");
  detail::CheckOp(input, rank2_op, target_axis, "ApplyRank2ToAxisPreserveOrder");
  if (ctx == nullptr) ctx = qlten::b200::detail::DefaultCtx();
  IndexVec<QNT> out_indexes = input.GetIndexes();
  out_indexes[target_axis] = rank2_op.GetIndex(1);
  if (!detail::RunAxis<ElemT, QNT>(input, &rank2_op, target_axis, nullptr, 0, out_indexes, out, ctx))
    detail::ViaContraction(input, rank2_op, target_axis, out, ctx);
}

template<typename ElemT, typename QNT>
void ApplyTwoRank2ToAxesPreserveOrder(const QLTensor<ElemT, QNT> &input, const QLTensor<ElemT, QNT> &op1, size_t axis1,
                                      const QLTensor<ElemT, QNT> &op2, size_t axis2, QLTensor<ElemT, QNT> &out, qlb200_ctx *ctx = nullptr) {
  static_assert(!Fermionicable<QNT>::IsFermionic(), "ApplyTwoRank2ToAxesPreserveOrder is bosonic-only.");
  if (&input == &out || &op1 == &out || &op2 == &out) throw std::invalid_argument("ApplyTwoRank2ToAxesPreserveOrder: output aliasing is not allowed.");
  if (axis1 == axis2) throw std::invalid_argument("ApplyTwoRank2ToAxesPreserveOrder: axes must be distinct.");
  detail::CheckOp(input, op1, axis1, "ApplyTwoRank2ToAxesPreserveOrder");
  detail::CheckOp(input, op2, axis2, "ApplyTwoRank2ToAxesPreserveOrder");
  if (ctx == nullptr) ctx = qlten::b200::detail::DefaultCtx();
  IndexVec<QNT> out_indexes = input.GetIndexes();
  out_indexes[axis1] = op1.GetIndex(1);
  out_indexes[axis2] = op2.GetIndex(1);
  if (!detail::RunAxis<ElemT, QNT>(input, &op1, axis1, &op2, axis2, out_indexes, out, ctx)) {
    QLTensor<ElemT, QNT> mid;
    detail::ViaContraction(input, op1, axis1, mid, ctx);
    detail::ViaContraction(mid, op2, axis2, out, ctx);
  }
}

}  // namespace dmrg
}  // namespace b200
}  // namespace qlten
#endif  // QLTEN_B200_AXIS_OPS_H
