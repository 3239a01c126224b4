// qlten_b200/contract.h -- drop-in adapter between TensorToolkit's QLTensor<ElemT, QNT> and the
// qlb200 C ABI (include/qlb200.h).
//
//   qlten::b200::Contract(pa, pb, axes_set, pc)               same signature/semantics as
//       qlten::Contract            (include/qlten/tensor_manipulation/ten_ctrct.h:277-290)
//   qlten::b200::Contract1Sector(pa, idx_a, sct, pb, axes, pc) as
//       qlten::dmrg::Contract1Sector (tensor_manipulation/dmrg/contract_1sector.h:211-228)
//   qlten::b200::Transpose(pt, order)                          as QLTensor::Transpose
//       (qltensor/qltensor_impl.h:449-464)
//   qlten::b200::ContractContiguousAxes<T, QNT, ASide, BSide>(a, b, a_start, b_start, size, c)  as
//       qlten::ContractContiguousAxes (tensor_manipulation/contract_contiguous_axes.h:849-873)
//   qlten::b200::ContractTailHeadContiguousAccumulate(a, b, a_start, b_start, size, alpha, beta, c, stats)  and
//   qlten::b200::TryContractTailHeadContiguousAccumulate(...)   as the reference's functions of the same names
//       (contract_contiguous_axes.h:954-1041):  c <- beta * c + alpha * contraction, output topology union included
//
// Only PUBLIC reference API is used (GetBlkSparDataTen, GetBlkIdxDataBlkMap, GetActualRawDataPtr,
// DataBlksInsert, TenCtrctGenSavedAxesSet, TenCtrctInitResTen), so the reference tree stays
// unmodified.  The QLTensor layout (one flat ElemT buffer, blocks in ascending blk_idx order,
// row-major inside a block) is kept as is; host tensors are staged through device memory by the
// library (QLB200_MEM_HOST).  Preconditions the reference only assert()s are checked and thrown.
#ifndef QLTEN_B200_CONTRACT_H
#define QLTEN_B200_CONTRACT_H

#include <complex>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "qlb200.h"
#include "qlten/qltensor_all.h"
#include "qlten/tensor_manipulation/ten_ctrct.h"
#include "qlten/tensor_manipulation/contract_contiguous_axes.h"

namespace qlten {
namespace b200 {

namespace detail {

inline void Check(int rc, const char *what) {
  if (rc != QLB200_OK) {
    throw std::runtime_error(std::string("qlb200: ") + what + ": " + qlb200_last_error());
  }
}

template<typename ElemT> struct DTypeOf;
template<> struct DTypeOf<QLTEN_Double> { static constexpr int value = QLB200_F64; };
template<> struct DTypeOf<QLTEN_Complex> { static constexpr int value = QLB200_C64; };

/// Owns the flat arrays a qlb200_shell points into.
struct ShellHolder {
  std::vector<uint32_t> nsct, deg, coors;
  std::vector<uint8_t> parity;
  std::vector<int8_t> dir;
  qlb200_shell shell{};
};

template<typename ElemT, typename QNT>
void FillShell(const QLTensor<ElemT, QNT> &t, ShellHolder &h) {
  const size_t rank = t.Rank();
  constexpr bool fermionic = Fermionicable<QNT>::IsFermionic();
  for (size_t i = 0; i < rank; ++i) {
    const Index<QNT> &idx = t.GetIndex(i);
    h.nsct.push_back(static_cast<uint32_t>(idx.GetQNSctNum()));
    h.dir.push_back(static_cast<int8_t>(idx.GetDir()));
    for (size_t s = 0; s < idx.GetQNSctNum(); ++s) {
      const QNSector<QNT> &sct = idx.GetQNSct(s);
      h.deg.push_back(static_cast<uint32_t>(sct.dim()));
      if constexpr (fermionic) { h.parity.push_back(sct.IsFermionParityOdd() ? 1 : 0); }
    }
  }
  const auto &blk_map = t.GetBlkSparDataTen().GetBlkIdxDataBlkMap();
  h.coors.reserve(blk_map.size() * rank);
  for (const auto &kv : blk_map) {
    for (size_t i = 0; i < rank; ++i) { h.coors.push_back(static_cast<uint32_t>(kv.second.blk_coors[i])); }
  }
  h.shell.rank = static_cast<int32_t>(rank);
  h.shell.nsct = h.nsct.data();
  h.shell.deg = h.deg.data();
  h.shell.parity = fermionic ? h.parity.data() : nullptr;
  h.shell.dir = h.dir.data();
  h.shell.nblk = blk_map.size();
  h.shell.blk_coors = h.coors.data();
}

struct MatchGuard {
  qlb200_match *m = nullptr;
  ~MatchGuard() { if (m) qlb200_match_destroy(m); }
};
struct PlanGuard {
  qlb200_plan *p = nullptr;
  ~PlanGuard() { if (p) qlb200_plan_destroy(p); }
};
struct TPlanGuard {
  qlb200_tplan *p = nullptr;
  ~TPlanGuard() { if (p) qlb200_tplan_destroy(p); }
};

/// Process-wide default context on device 0 (the reference keeps process-wide singletons for
/// its cuBLAS/cuTENSOR handles too: framework/hp_numeric/gpu_set.h:30-125).
inline qlb200_ctx *DefaultCtx() {
  static qlb200_ctx *ctx = [] {
    qlb200_ctx *c = nullptr;
    Check(qlb200_ctx_create(0, &c), "ctx_create");
    return c;
  }();
  return ctx;
}

template<typename ElemT, typename QNT>
void CheckPreconditions(const QLTensor<ElemT, QNT> *pa, const QLTensor<ElemT, QNT> *pb,
                        const std::vector<std::vector<size_t>> &axes_set,
                        const QLTensor<ElemT, QNT> *pc) {
  if (!pc->IsDefault()) { throw std::invalid_argument("b200::Contract: result tensor must be default"); }
  if (axes_set.size() != 2 || axes_set[0].size() != axes_set[1].size()) {
    throw std::invalid_argument("b200::Contract: malformed axes_set");
  }
  for (size_t i = 0; i < axes_set[0].size(); ++i) {
    if (axes_set[0][i] >= pa->Rank() || axes_set[1][i] >= pb->Rank() ||
        !(pa->GetIndex(axes_set[0][i]) == InverseIndex(pb->GetIndex(axes_set[1][i])))) {
      throw std::invalid_argument("b200::Contract: contracted indexes do not match");
    }
  }
}

template<typename ElemT, typename QNT>
void RunMatched(const QLTensor<ElemT, QNT> *pa, const QLTensor<ElemT, QNT> *pb,
                const ShellHolder &sa, const ShellHolder &sb, qlb200_match *m,
                QLTensor<ElemT, QNT> *pc, qlb200_ctx *ctx) {
  const uint64_t ntask = qlb200_match_ntask(m);
  if (ntask == 0) { return; }   // reference: CtrctTwoBSDTAndAssignIn returns early, C keeps no data
  PlanGuard plan;
  Check(qlb200_plan_create(ctx, m, &sa.shell, &sb.shell, DTypeOf<ElemT>::value,
                           QLB200_PLAN_DETERMINISTIC, &plan.p), "plan_create");
  const ElemT *a_raw = pa->GetBlkSparDataTen().GetActualRawDataPtr();
  const ElemT *b_raw = pb->GetBlkSparDataTen().GetActualRawDataPtr();
  if (qlb200_match_is_scalar(m)) {
    ElemT v(0);
    Check(qlb200_execute(ctx, plan.p, a_raw, b_raw, &v, QLB200_MEM_HOST), "execute");
    pc->SetElem({}, v);
    return;
  }
  const int32_t c_rank = qlb200_match_c_rank(m);
  const uint64_t c_nblk = qlb200_match_c_nblk(m);
  std::vector<uint64_t> blk_idx(c_nblk), off(c_nblk);
  std::vector<uint32_t> coors(c_nblk * c_rank), shape(c_nblk * c_rank);
  Check(qlb200_match_c_blocks(m, blk_idx.data(), coors.data(), shape.data(), off.data()), "c_blocks");
  std::vector<size_t> idxs(blk_idx.begin(), blk_idx.end());
  std::vector<CoorsT> coors_s(c_nblk, CoorsT(c_rank));
  for (uint64_t b = 0; b < c_nblk; ++b) {
    for (int32_t i = 0; i < c_rank; ++i) { coors_s[b][i] = coors[b * c_rank + i]; }
  }
  auto &bsdt_c = pc->GetBlkSparDataTen();
  bsdt_c.DataBlksInsert(idxs, coors_s, true);   // sets offsets + raw_data_size_, allocates (uninitialised)
  Check(qlb200_execute(ctx, plan.p, a_raw, b_raw, bsdt_c.GetActualRawDataPtr(), QLB200_MEM_HOST), "execute");
}

}  // namespace detail

template<typename TenElemT, typename QNT>
void Contract(const QLTensor<TenElemT, QNT> *pa, const QLTensor<TenElemT, QNT> *pb,
              const std::vector<std::vector<size_t>> &axes_set, QLTensor<TenElemT, QNT> *pc,
              qlb200_ctx *ctx = nullptr) {
  detail::CheckPreconditions(pa, pb, axes_set, pc);
  if (ctx == nullptr) { ctx = detail::DefaultCtx(); }
  auto saved_axes_set = TenCtrctGenSavedAxesSet(pa->Rank(), pb->Rank(), axes_set);
  TenCtrctInitResTen(pa, pb, saved_axes_set, pc);
  detail::ShellHolder sa, sb;
  detail::FillShell(*pa, sa);
  detail::FillShell(*pb, sb);
  std::vector<int32_t> aa(axes_set[0].begin(), axes_set[0].end()), ba(axes_set[1].begin(), axes_set[1].end());
  detail::MatchGuard match;
  detail::Check(qlb200_match_create(&sa.shell, &sb.shell, static_cast<int32_t>(aa.size()), aa.data(),
                                    ba.data(), &match.m), "match_create");
  detail::RunMatched(pa, pb, sa, sb, match.m, pc, ctx);
}

/// Mixed real/complex overloads promote like the reference does (ten_ctrct.h:292-350).
template<typename QNT>
void Contract(const QLTensor<QLTEN_Double, QNT> *pa, const QLTensor<QLTEN_Complex, QNT> *pb,
              const std::vector<std::vector<size_t>> &axes_set, QLTensor<QLTEN_Complex, QNT> *pc,
              qlb200_ctx *ctx = nullptr) {
  auto cplx_a = ToComplex(*pa);
  Contract(&cplx_a, pb, axes_set, pc, ctx);
}
template<typename QNT>
void Contract(const QLTensor<QLTEN_Complex, QNT> *pa, const QLTensor<QLTEN_Double, QNT> *pb,
              const std::vector<std::vector<size_t>> &axes_set, QLTensor<QLTEN_Complex, QNT> *pc,
              qlb200_ctx *ctx = nullptr) {
  auto cplx_b = ToComplex(*pb);
  Contract(pa, &cplx_b, axes_set, pc, ctx);
}

template<typename TenElemT, typename QNT>
void Contract1Sector(const QLTensor<TenElemT, QNT> *pa, const size_t idx_a, const size_t qn_sector_idx_a,
                     const QLTensor<TenElemT, QNT> *pb, const std::vector<std::vector<size_t>> &axes_set,
                     QLTensor<TenElemT, QNT> *pc, qlb200_ctx *ctx = nullptr) {
  detail::CheckPreconditions(pa, pb, axes_set, pc);
  if (idx_a >= pa->Rank() || qn_sector_idx_a >= pa->GetIndex(idx_a).GetQNSctNum()) {
    throw std::invalid_argument("b200::Contract1Sector: bad split index / sector");
  }
  for (size_t ax : axes_set[0]) {
    if (ax == idx_a) { throw std::invalid_argument("b200::Contract1Sector: split index is contracted"); }
  }
  if (ctx == nullptr) { ctx = detail::DefaultCtx(); }
  auto saved_axes_set = TenCtrctGenSavedAxesSet(pa->Rank(), pb->Rank(), axes_set);
  TenCtrctInitResTen(pa, pb, saved_axes_set, pc);
  detail::ShellHolder sa, sb;
  detail::FillShell(*pa, sa);
  detail::FillShell(*pb, sb);
  std::vector<int32_t> aa(axes_set[0].begin(), axes_set[0].end()), ba(axes_set[1].begin(), axes_set[1].end());
  detail::MatchGuard match;
  detail::Check(qlb200_match_create_1sector(&sa.shell, static_cast<int32_t>(idx_a),
                                            static_cast<uint32_t>(qn_sector_idx_a), &sb.shell,
                                            static_cast<int32_t>(aa.size()), aa.data(), ba.data(), &match.m),
                "match_create_1sector");
  detail::RunMatched(pa, pb, sa, sb, match.m, pc, ctx);
}

/// Same signature and result as qlten::ContractContiguousAxes.  ASide / BSide only tell the reference which operand
/// to transpose physically; here every block is read in place by the grouped GEMM whatever the sides are, so they
/// are accepted and ignored.
template<typename TenElemT, typename QNT, CtrctSide ASide = CtrctSide::Tail, CtrctSide BSide = CtrctSide::Head>
void ContractContiguousAxes(const QLTensor<TenElemT, QNT> &a, const QLTensor<TenElemT, QNT> &b,
                            const size_t a_ctrct_axes_start, const size_t b_ctrct_axes_start,
                            const size_t ctrct_axes_size, QLTensor<TenElemT, QNT> &c, qlb200_ctx *ctx = nullptr) {
  const size_t ra = a.Rank(), rb = b.Rank();
  if (ra == 0 || rb == 0 || a_ctrct_axes_start >= ra || b_ctrct_axes_start >= rb || ctrct_axes_size > ra ||
      ctrct_axes_size > rb) {
    throw std::invalid_argument("b200::ContractContiguousAxes: bad axis range");
  }
  std::vector<std::vector<size_t>> axes_set(2), saved_axes_set(2);
  for (size_t i = 0; i < ctrct_axes_size; ++i) {
    axes_set[0].push_back((a_ctrct_axes_start + i) % ra);
    axes_set[1].push_back((b_ctrct_axes_start + i) % rb);
  }
  for (size_t i = 0; i < ra - ctrct_axes_size; ++i) { saved_axes_set[0].push_back((a_ctrct_axes_start + ctrct_axes_size + i) % ra); }
  for (size_t i = 0; i < rb - ctrct_axes_size; ++i) { saved_axes_set[1].push_back((b_ctrct_axes_start + ctrct_axes_size + i) % rb); }
  c = QLTensor<TenElemT, QNT>();     // the reference re-initialises the output (TenCtrctInitResTen on whatever c held)
  detail::CheckPreconditions(&a, &b, axes_set, &c);
  if (ctx == nullptr) { ctx = detail::DefaultCtx(); }
  TenCtrctInitResTen(&a, &b, saved_axes_set, &c);
  detail::ShellHolder sa, sb;
  detail::FillShell(a, sa);
  detail::FillShell(b, sb);
  detail::MatchGuard match;
  detail::Check(qlb200_match_create_contiguous(&sa.shell, &sb.shell, static_cast<int32_t>(a_ctrct_axes_start),
                                               static_cast<int32_t>(b_ctrct_axes_start),
                                               static_cast<int32_t>(ctrct_axes_size), &match.m),
                "match_create_contiguous");
  detail::RunMatched(&a, &b, sa, sb, match.m, &c, ctx);
}

namespace detail {

struct AccumGuard {
  qlb200_accum *a = nullptr;
  ~AccumGuard() { if (a) qlb200_accum_destroy(a); }
};

/// The reference's own exception type for layout mismatches, so that callers (and Try...) can tell it apart.
using LayoutMismatch = qlten::detail::ContractAccumulateLayoutMismatch;

template<typename ElemT> inline void SplitScalar(const ElemT &v, double out[2]);
template<> inline void SplitScalar<QLTEN_Double>(const QLTEN_Double &v, double out[2]) { out[0] = v; out[1] = 0.0; }
template<> inline void SplitScalar<QLTEN_Complex>(const QLTEN_Complex &v, double out[2]) { out[0] = v.real(); out[1] = v.imag(); }

template<typename TenElemT, typename QNT>
void ContractAccumulateImpl(const QLTensor<TenElemT, QNT> &a, const QLTensor<TenElemT, QNT> &b, const size_t a_start,
                            const size_t b_start, const size_t size, const TenElemT alpha, const TenElemT beta,
                            QLTensor<TenElemT, QNT> &c, ContiguousContractStats *stats, const bool allow_expand, qlb200_ctx *ctx) {
  const size_t ra = a.Rank(), rb = b.Rank();
  if (ra == 0 || rb == 0 || a_start >= ra || b_start >= rb || size > ra || size > rb) {
    throw std::invalid_argument("b200::ContractTailHeadContiguousAccumulate: bad axis range");
  }
  std::vector<std::vector<size_t>> axes_set(2), saved_axes_set(2);
  for (size_t i = 0; i < size; ++i) { axes_set[0].push_back((a_start + i) % ra); axes_set[1].push_back((b_start + i) % rb); }
  for (size_t i = 0; i < ra - size; ++i) { saved_axes_set[0].push_back((a_start + size + i) % ra); }
  for (size_t i = 0; i < rb - size; ++i) { saved_axes_set[1].push_back((b_start + size + i) % rb); }
  for (size_t i = 0; i < size; ++i) {
    if (!(a.GetIndex(axes_set[0][i]) == InverseIndex(b.GetIndex(axes_set[1][i])))) {
      throw std::invalid_argument("b200::ContractTailHeadContiguousAccumulate: contracted indexes do not match");
    }
  }
  const bool c_default = c.IsDefault();
  if (c_default && beta != TenElemT(0)) {
    throw std::invalid_argument("ContractTailHeadContiguousAccumulate requires beta == 0 when output is default.");
  }
  const auto expected_indexes = qlten::detail::MakeContractResultIndexes(a, b, saved_axes_set);
  if (!c_default && c.GetIndexes() != expected_indexes) {
    throw LayoutMismatch("ContractTailHeadContiguousAccumulate output indexes are not compatible with the contraction result.");
  }
  if (ctx == nullptr) { ctx = DefaultCtx(); }
  ShellHolder sa, sb, sc;
  FillShell(a, sa);
  FillShell(b, sb);
  MatchGuard match;
  Check(qlb200_match_create_contiguous(&sa.shell, &sb.shell, static_cast<int32_t>(a_start), static_cast<int32_t>(b_start),
                                       static_cast<int32_t>(size), &match.m), "match_create_contiguous");
  const bool has_data = !c_default && c.GetBlkSparDataTen().GetActualRawDataSize() > 0;
  if (!c_default) { FillShell(c, sc); }
  double al[2], be[2];
  SplitScalar(alpha, al);
  SplitScalar(beta, be);
  AccumGuard acc;
  const int rc = qlb200_accum_create(match.m, c_default ? nullptr : &sc.shell, has_data ? 1 : 0, allow_expand ? 1 : 0,
                                     DTypeOf<TenElemT>::value, al, be, &acc.a);
  if (rc == QLB200_ERR_LAYOUT) { throw LayoutMismatch(std::string("ContractTailHeadContiguousAccumulate ") + qlb200_last_error()); }
  if (rc == QLB200_ERR_ARG) { throw std::invalid_argument(std::string("ContractTailHeadContiguousAccumulate ") + qlb200_last_error()); }
  Check(rc, "accum_create");
  if (stats != nullptr) {
    qlb200_accum_stats st;
    Check(qlb200_accum_get_stats(acc.a, &st), "accum_get_stats");
    stats->raw_data_contract_tasks = st.raw_data_contract_tasks;
    stats->gemm_calls = st.gemm_calls;
    stats->accumulate_calls = st.accumulate_calls;
    stats->accumulate_gemm_calls = st.accumulate_gemm_calls;
    stats->output_tensor_rebuilds = st.output_tensor_rebuilds;
    stats->temporary_output_bytes_avoided = st.temporary_output_bytes_avoided;
    stats->output_topology_expansions = st.output_topology_expansions;
    stats->output_expand_copy_bytes = st.output_expand_copy_bytes;
    stats->output_expand_new_blocks = st.output_expand_new_blocks;
    stats->output_untouched_scale_bytes = st.output_untouched_scale_bytes;
  }
  const uint64_t ntask = qlb200_match_ntask(match.m);
  const bool scalar = qlb200_match_is_scalar(match.m) != 0;
  const bool expanded = qlb200_accum_expanded(acc.a) != 0;
  // the output tensor on the resulting topology: a fresh one when c was default or must grow, c itself otherwise
  QLTensor<TenElemT, QNT> grown;
  QLTensor<TenElemT, QNT> *out = &c;
  const TenElemT *old_raw = has_data ? c.GetBlkSparDataTen().GetActualRawDataPtr() : nullptr;
  if (c_default || expanded) {
    grown = QLTensor<TenElemT, QNT>(expected_indexes);
    out = &grown;
  }
  if (scalar) {
    if (!out->IsScalar() || out->GetBlkSparDataTen().GetActualRawDataSize() == 0) { out->SetElem({}, TenElemT(0)); }
  } else if (c_default || expanded) {
    const uint64_t nblk = qlb200_accum_nblk(acc.a);
    const int32_t c_rank = qlb200_match_c_rank(match.m);
    std::vector<uint64_t> blk_idx(nblk);
    std::vector<uint32_t> coors(nblk * c_rank);
    Check(qlb200_accum_blocks(acc.a, blk_idx.data(), coors.data(), nullptr, nullptr, nullptr, nullptr), "accum_blocks");
    std::vector<size_t> idxs(blk_idx.begin(), blk_idx.end());
    std::vector<CoorsT> coors_s(nblk, CoorsT(c_rank));
    for (uint64_t blk = 0; blk < nblk; ++blk) {
      for (int32_t i = 0; i < c_rank; ++i) { coors_s[blk][i] = coors[blk * c_rank + i]; }
    }
    if (nblk > 0) { out->GetBlkSparDataTen().DataBlksInsert(idxs, coors_s, true); }
  } else if (!has_data) {
    out->GetBlkSparDataTen().Allocate(true);     // existing shell without raw data (beta == 0 checked by the library)
  }
  TenElemT *new_raw = out->GetBlkSparDataTen().GetActualRawDataPtr();
  if (new_raw != nullptr && (ntask > 0 || !c_default)) {
    PlanGuard plan;
    Check(qlb200_plan_create_accum(ctx, match.m, acc.a, DTypeOf<TenElemT>::value, QLB200_PLAN_DETERMINISTIC, &plan.p), "plan_create_accum");
    Check(qlb200_execute_accum(ctx, plan.p, a.GetBlkSparDataTen().GetActualRawDataPtr(), b.GetBlkSparDataTen().GetActualRawDataPtr(),
                               old_raw, new_raw, QLB200_MEM_HOST), "execute_accum");
  }
  if (out != &c) { c = std::move(grown); }
}

}  // namespace detail

/// c <- beta * c + alpha * ContractTailHeadContiguous(a, b, ...): same signature and semantics as
/// qlten::ContractTailHeadContiguousAccumulate (contract_contiguous_axes.h:954-1000), incl. the aliasing check, the
/// beta == 0 rule for a default output, output blocks beyond the contraction's (scaled by beta only), output-topology
/// expansion to the union of old and required blocks, and the ContiguousContractStats counters.
template<typename TenElemT, typename QNT>
void ContractTailHeadContiguousAccumulate(const QLTensor<TenElemT, QNT> &pa, const QLTensor<TenElemT, QNT> &pb,
                                          const size_t a_ctrct_axes_start, const size_t b_ctrct_axes_start,
                                          const size_t ctrct_axes_size, const TenElemT alpha, const TenElemT beta,
                                          QLTensor<TenElemT, QNT> &pc, ContiguousContractStats *stats = nullptr,
                                          qlb200_ctx *ctx = nullptr) {
  if (stats != nullptr) { *stats = ContiguousContractStats{}; }
  if (&pc == &pa || &pc == &pb) {
    throw std::invalid_argument("ContractTailHeadContiguousAccumulate does not support aliasing between output and input tensors.");
  }
  detail::ContractAccumulateImpl(pa, pb, a_ctrct_axes_start, b_ctrct_axes_start, ctrct_axes_size, alpha, beta, pc, stats, true, ctx);
}

/// The no-expansion probe (contract_contiguous_axes.h:1002-1041): false -- pc untouched, stats reset -- on a layout mismatch.
template<typename TenElemT, typename QNT>
bool TryContractTailHeadContiguousAccumulate(const QLTensor<TenElemT, QNT> &pa, const QLTensor<TenElemT, QNT> &pb,
                                             const size_t a_ctrct_axes_start, const size_t b_ctrct_axes_start,
                                             const size_t ctrct_axes_size, const TenElemT alpha, const TenElemT beta,
                                             QLTensor<TenElemT, QNT> &pc, ContiguousContractStats *stats = nullptr,
                                             qlb200_ctx *ctx = nullptr) {
  if (stats != nullptr) { *stats = ContiguousContractStats{}; }
  if (&pc == &pa || &pc == &pb) {
    throw std::invalid_argument("TryContractTailHeadContiguousAccumulate does not support aliasing between output and input tensors.");
  }
  try {
    detail::ContractAccumulateImpl(pa, pb, a_ctrct_axes_start, b_ctrct_axes_start, ctrct_axes_size, alpha, beta, pc, stats, false, ctx);
    return true;
  } catch (const detail::LayoutMismatch &) {
    if (stats != nullptr) { *stats = ContiguousContractStats{}; }
    return false;
  }
}

template<typename TenElemT, typename QNT>
void Transpose(QLTensor<TenElemT, QNT> *pt, const std::vector<size_t> &transed_idxes_order,
               qlb200_ctx *ctx = nullptr) {
  if (pt->IsDefault()) { throw std::invalid_argument("b200::Transpose: default tensor"); }
  if (pt->IsScalar()) { return; }
  if (transed_idxes_order.size() != pt->Rank()) { throw std::invalid_argument("b200::Transpose: bad order"); }
  if (std::is_sorted(transed_idxes_order.begin(), transed_idxes_order.end())) { return; }
  if (ctx == nullptr) { ctx = detail::DefaultCtx(); }
  const size_t rank = pt->Rank();
  detail::ShellHolder st;
  detail::FillShell(*pt, st);
  std::vector<int32_t> perm(transed_idxes_order.begin(), transed_idxes_order.end());
  detail::TPlanGuard tp;
  detail::Check(qlb200_tplan_create(ctx, &st.shell, perm.data(), detail::DTypeOf<TenElemT>::value, &tp.p),
                "tplan_create");
  IndexVec<QNT> new_idxs;
  for (size_t i = 0; i < rank; ++i) { new_idxs.push_back(pt->GetIndex(transed_idxes_order[i])); }
  QLTensor<TenElemT, QNT> out(new_idxs);
  const uint64_t nblk = qlb200_tplan_nblk(tp.p);
  if (nblk > 0) {
    std::vector<uint64_t> blk_idx(nblk), off(nblk);
    std::vector<uint32_t> coors(nblk * rank), shape(nblk * rank);
    std::vector<int8_t> scale(nblk);
    detail::Check(qlb200_tplan_blocks(tp.p, blk_idx.data(), coors.data(), shape.data(), off.data(), scale.data()),
                  "tplan_blocks");
    std::vector<size_t> idxs(blk_idx.begin(), blk_idx.end());
    std::vector<CoorsT> coors_s(nblk, CoorsT(rank));
    for (uint64_t b = 0; b < nblk; ++b) {
      for (size_t i = 0; i < rank; ++i) { coors_s[b][i] = coors[b * rank + i]; }
    }
    out.GetBlkSparDataTen().DataBlksInsert(idxs, coors_s, true);
    detail::Check(qlb200_transpose_execute(ctx, tp.p, pt->GetBlkSparDataTen().GetActualRawDataPtr(),
                                           out.GetBlkSparDataTen().GetActualRawDataPtr(), QLB200_MEM_HOST),
                  "transpose_execute");
  }
  *pt = std::move(out);
}

}  // namespace b200
}  // namespace qlten
#endif  // QLTEN_B200_CONTRACT_H
