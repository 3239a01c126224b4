// oracle/ref_handles.h -- TEST INFRASTRUCTURE ONLY.
//
// Opaque-handle types shared by the two oracle libraries:
//   oracle/_ref/libqlref.so      (ref_driver.cc)      the UNMODIFIED reference behind a C ABI -- links nothing of the product;
//   oracle/_ref/libqladapter.so  (adapter_driver.cc)  the drop-in adapter include/qlten_b200/*.h instantiated on the same
//                                                     reference tensor types -- links tensortoolkit_b200/libqlb200.so.
// A handle created by one library is a TenBase* the other can read: both compile this header.
#ifndef QLREF_HANDLES_H
#define QLREF_HANDLES_H

#include <cstdint>
#include <cstring>
#include <fstream>
#include <chrono>
#include <algorithm>
#include <memory>
#include <unordered_map>
#include <string>
#include <vector>

#include "qlten/qltensor_all.h"
#include "qlten/tensor_manipulation/ten_ctrct.h"
#include "qlten/tensor_manipulation/tensor_op_cost.h"
#include "qlten/tensor_manipulation/dmrg/contract_1sector.h"
#include "qlten/tensor_manipulation/contract_contiguous_axes.h"
#include "qlten/tensor_manipulation/dmrg/axis_ops.h"

using namespace qlten;
using special_qn::U1QN;
using special_qn::U1U1QN;
using special_qn::fU1QN;
using special_qn::fU1U1QN;
using special_qn::Z2QN;
using special_qn::fZ2QN;

namespace qlref {

enum QNKind { K_U1 = 0, K_FU1 = 1, K_U1U1 = 2, K_FU1U1 = 3, K_Z2 = 4, K_FZ2 = 5 };
enum DType { D_F64 = 0, D_C64 = 1 };

template<typename QNT> struct QNMake;
template<> struct QNMake<U1QN> { static U1QN make(const int64_t *v) { return U1QN((int) v[0]); } static constexpr int nv = 1; };
template<> struct QNMake<fU1QN> { static fU1QN make(const int64_t *v) { return fU1QN((int) v[0]); } static constexpr int nv = 1; };
template<> struct QNMake<U1U1QN> { static U1U1QN make(const int64_t *v) { return U1U1QN((int) v[0], (int) v[1]); } static constexpr int nv = 2; };
template<> struct QNMake<fU1U1QN> { static fU1U1QN make(const int64_t *v) { return fU1U1QN((int) v[0], (int) v[1]); } static constexpr int nv = 2; };
template<> struct QNMake<Z2QN> { static Z2QN make(const int64_t *v) { return Z2QN((int) v[0]); } static constexpr int nv = 1; };
template<> struct QNMake<fZ2QN> { static fZ2QN make(const int64_t *v) { return fZ2QN((int) v[0]); } static constexpr int nv = 1; };

struct TaskRec {  // mirrors qlb200_task field meaning, filled from RawDataCtrctTask
  uint64_t a_blk_idx, b_blk_idx, c_blk_idx, a_off, b_off, c_off;
  uint64_t m, k, n;
  double sign, beta;
};

struct IdxBase {
  virtual ~IdxBase() {}
  int kind;
};
template<typename QNT> struct IdxBox : IdxBase { Index<QNT> idx; };

struct TenBase {
  virtual ~TenBase() {}
  int kind, dtype;
  virtual int rank() const = 0;
  virtual bool is_default() const = 0;
  virtual uint64_t nblk() const = 0;
  virtual uint64_t raw_size() const = 0;
  virtual const void *raw() const = 0;
  virtual void *raw_mut() = 0;
  virtual void nsct(uint32_t *out) const = 0;
  virtual void dirs(int8_t *out) const = 0;
  virtual void degs(uint32_t *out) const = 0;      // concatenated over indexes
  virtual void parities(uint8_t *out) const = 0;   // concatenated over indexes (0 for bosonic)
  virtual bool fermionic() const = 0;
  virtual void blocks(uint64_t *idx, uint32_t *coors, uint32_t *shape, uint64_t *off) const = 0;
  virtual void random(const int64_t *div) = 0;
  virtual TenBase *clone() const = 0;
  virtual void reset_default() = 0;
  virtual void transpose(const int64_t *perm) = 0;
  virtual double norm2() const = 0;
  virtual bool indexes_equal(const TenBase *o) const = 0;
  virtual TenBase *contract(const TenBase *b, int n, const int64_t *aa, const int64_t *ba) const = 0;
  virtual TenBase *contract_1sector(int64_t axis, int64_t sct, const TenBase *b, int n, const int64_t *aa, const int64_t *ba) const = 0;
  virtual int write_file(const char *path) const = 0;
  virtual int read_file(const char *path) = 0;
  // side: 0 = <Tail, Head> (default), 1 = <Head, Head>, 2 = <Tail, Tail>, 3 = <Head, Tail>
  virtual TenBase *contract_contiguous(const TenBase *b, int64_t a_start, int64_t b_start, int64_t size, int side) const = 0;
  // c <- beta * c + alpha * contraction (ContractTailHeadContiguousAccumulate; try_only: the Try... probe).  Returns 1, or 0
  // when the probe reports a layout mismatch; stats10 = the ContiguousContractStats counters this path defines.
  virtual int contract_accumulate(const TenBase *b, int64_t a_start, int64_t b_start, int64_t size, const double *alpha2,
                                  const double *beta2, TenBase *c, int try_only, uint64_t *stats10) const = 0;
  // dmrg::ApplyRank2ToAxisPreserveOrder (op2 == nullptr) / ApplyTwoRank2ToAxesPreserveOrder; bosonic kinds only (nullptr otherwise)
  virtual TenBase *apply_rank2(const TenBase *op1, int64_t axis1, const TenBase *op2, int64_t axis2) const = 0;
  virtual std::vector<TaskRec> tasks(const TenBase *b, int n, const int64_t *aa, const int64_t *ba, bool sorted) const = 0;
  virtual void cost(const TenBase *b, int n, const int64_t *aa, const int64_t *ba, double *out8) const = 0;
};

inline std::vector<std::vector<size_t>> MakeAxes(int n, const int64_t *aa, const int64_t *ba) {
  std::vector<std::vector<size_t>> axes(2);
  for (int i = 0; i < n; ++i) { axes[0].push_back((size_t) aa[i]); axes[1].push_back((size_t) ba[i]); }
  return axes;
}

template<typename ElemT, typename QNT>
struct TenBox : TenBase {
  using Ten = QLTensor<ElemT, QNT>;
  using Elem = ElemT;
  using QN = QNT;
  Ten t;
  static const TenBox *cast(const TenBase *b) { return static_cast<const TenBox *>(b); }
  int rank() const override { return (int) t.Rank(); }
  bool is_default() const override { return t.IsDefault(); }
  uint64_t nblk() const override { return t.IsDefault() ? 0 : t.GetBlkSparDataTen().GetBlkIdxDataBlkMap().size(); }
  uint64_t raw_size() const override { return t.IsDefault() ? 0 : t.GetBlkSparDataTen().GetActualRawDataSize(); }
  const void *raw() const override { return t.IsDefault() ? nullptr : t.GetBlkSparDataTen().GetActualRawDataPtr(); }
  void *raw_mut() override { return t.IsDefault() ? nullptr : t.GetBlkSparDataTen().GetActualRawDataPtr(); }
  void nsct(uint32_t *out) const override { for (size_t i = 0; i < t.Rank(); ++i) out[i] = (uint32_t) t.GetIndex(i).GetQNSctNum(); }
  void dirs(int8_t *out) const override { for (size_t i = 0; i < t.Rank(); ++i) out[i] = (int8_t) t.GetIndex(i).GetDir(); }
  void degs(uint32_t *out) const override {
    size_t p = 0;
    for (size_t i = 0; i < t.Rank(); ++i)
      for (size_t s = 0; s < t.GetIndex(i).GetQNSctNum(); ++s) out[p++] = (uint32_t) t.GetIndex(i).GetQNSct(s).GetDegeneracy();
  }
  bool fermionic() const override { return Fermionicable<QNT>::IsFermionic(); }
  void parities(uint8_t *out) const override {
    size_t p = 0;
    for (size_t i = 0; i < t.Rank(); ++i)
      for (size_t s = 0; s < t.GetIndex(i).GetQNSctNum(); ++s) {
        if constexpr (Fermionicable<QNT>::IsFermionic()) out[p++] = t.GetIndex(i).GetQNSct(s).IsFermionParityOdd() ? 1 : 0;
        else out[p++] = 0;
      }
  }
  void blocks(uint64_t *idx, uint32_t *coors, uint32_t *shape, uint64_t *off) const override {
    if (t.IsDefault()) return;
    size_t r = t.Rank(), b = 0;
    for (const auto &[bi, blk] : t.GetBlkSparDataTen().GetBlkIdxDataBlkMap()) {
      idx[b] = bi; off[b] = blk.data_offset;
      for (size_t i = 0; i < r; ++i) { coors[b * r + i] = (uint32_t) blk.blk_coors[i]; shape[b * r + i] = (uint32_t) blk.shape[i]; }
      ++b;
    }
  }
  void random(const int64_t *div) override { t.Random(QNMake<QNT>::make(div)); }
  TenBase *clone() const override { auto *p = new TenBox(*this); return p; }
  void reset_default() override { t = Ten(); }
  void transpose(const int64_t *perm) override {
    std::vector<size_t> o(t.Rank()); for (size_t i = 0; i < t.Rank(); ++i) o[i] = (size_t) perm[i];
    t.Transpose(o);
  }
  double norm2() const override { return (double) t.Get2Norm(); }
  int write_file(const char *path) const override {          // the reference's own stream format (operator<<)
    std::ofstream ofs(path, std::ofstream::binary);
    if (!ofs) return -1;
    ofs << t;
    return ofs ? 0 : -1;
  }
  int read_file(const char *path) override {                 // operator>> into a default tensor
    std::ifstream ifs(path, std::ifstream::binary);
    if (!ifs) return -1;
    t = Ten();
    ifs >> t;
    return ifs ? 0 : -1;
  }
  bool indexes_equal(const TenBase *o) const override { return t.GetIndexes() == cast(o)->t.GetIndexes(); }
  TenBase *wrap(Ten &&c) const { auto *p = new TenBox(); p->kind = kind; p->dtype = dtype; p->t = std::move(c); return p; }
  TenBase *contract(const TenBase *b, int n, const int64_t *aa, const int64_t *ba) const override {
    Ten c; Contract(&t, &cast(b)->t, MakeAxes(n, aa, ba), &c); return wrap(std::move(c));
  }
  TenBase *contract_1sector(int64_t axis, int64_t sct, const TenBase *b, int n, const int64_t *aa, const int64_t *ba) const override {
    Ten c; dmrg::Contract1Sector(&t, (size_t) axis, (size_t) sct, &cast(b)->t, MakeAxes(n, aa, ba), &c); return wrap(std::move(c));
  }
  TenBase *contract_contiguous(const TenBase *b, int64_t a_start, int64_t b_start, int64_t size, int side) const override {
    Ten c;
    const Ten &tb = cast(b)->t;
    switch (side) {
      case 1: ContractContiguousAxes<ElemT, QNT, CtrctSide::Head, CtrctSide::Head>(t, tb, (size_t) a_start, (size_t) b_start, (size_t) size, c); break;
      case 2: ContractContiguousAxes<ElemT, QNT, CtrctSide::Tail, CtrctSide::Tail>(t, tb, (size_t) a_start, (size_t) b_start, (size_t) size, c); break;
      case 3: ContractContiguousAxes<ElemT, QNT, CtrctSide::Head, CtrctSide::Tail>(t, tb, (size_t) a_start, (size_t) b_start, (size_t) size, c); break;
      default: ContractContiguousAxes<ElemT, QNT, CtrctSide::Tail, CtrctSide::Head>(t, tb, (size_t) a_start, (size_t) b_start, (size_t) size, c); break;
    }
    return wrap(std::move(c));
  }
  TenBase *apply_rank2(const TenBase *op1, int64_t axis1, const TenBase *op2, int64_t axis2) const override {
    if constexpr (Fermionicable<QNT>::IsFermionic()) {
      return nullptr;
    } else {
      Ten out;
      if (op2 == nullptr) dmrg::ApplyRank2ToAxisPreserveOrder(t, cast(op1)->t, (size_t) axis1, out);
      else dmrg::ApplyTwoRank2ToAxesPreserveOrder(t, cast(op1)->t, (size_t) axis1, cast(op2)->t, (size_t) axis2, out);
      return wrap(std::move(out));
    }
  }
  static ElemT MakeScalar(const double *v2) {
    if constexpr (std::is_same<ElemT, QLTEN_Complex>::value) return ElemT(v2[0], v2[1]);
    else return ElemT(v2[0]);
  }
  static void StoreStats(const ContiguousContractStats &st, uint64_t *o) {
    o[0] = st.raw_data_contract_tasks; o[1] = st.gemm_calls; o[2] = st.accumulate_calls; o[3] = st.accumulate_gemm_calls;
    o[4] = st.output_tensor_rebuilds; o[5] = st.temporary_output_bytes_avoided; o[6] = st.output_topology_expansions;
    o[7] = st.output_expand_copy_bytes; o[8] = st.output_expand_new_blocks; o[9] = st.output_untouched_scale_bytes;
  }
  int contract_accumulate(const TenBase *b, int64_t a_start, int64_t b_start, int64_t size, const double *alpha2, const double *beta2,
                          TenBase *c, int try_only, uint64_t *stats10) const override {
    ContiguousContractStats st;
    Ten &tc = static_cast<TenBox *>(c)->t;
    int ok = 1;
    if (try_only) {
      ok = TryContractTailHeadContiguousAccumulate(t, cast(b)->t, (size_t) a_start, (size_t) b_start, (size_t) size, MakeScalar(alpha2),
                                                   MakeScalar(beta2), tc, &st) ? 1 : 0;
    } else {
      ContractTailHeadContiguousAccumulate(t, cast(b)->t, (size_t) a_start, (size_t) b_start, (size_t) size, MakeScalar(alpha2),
                                           MakeScalar(beta2), tc, &st);
    }
    if (stats10) StoreStats(st, stats10);
    return ok;
  }
  std::vector<TaskRec> tasks(const TenBase *b, int n, const int64_t *aa, const int64_t *ba, bool sorted) const override {
    auto axes = MakeAxes(n, aa, ba);
    auto saved = TenCtrctGenSavedAxesSet(t.Rank(), cast(b)->t.Rank(), axes);
    Ten c; TenCtrctInitResTen(&t, &cast(b)->t, saved, &c);
    auto rt = c.GetBlkSparDataTen().DataBlkGenForTenCtrct(t.GetBlkSparDataTen(), cast(b)->t.GetBlkSparDataTen(), axes, saved);
    if (sorted) RawDataCtrctTask::SortTasksByCBlkIdx(rt);
    std::vector<TaskRec> out;
    for (auto &x : rt) out.push_back({x.a_blk_idx, x.b_blk_idx, x.c_blk_idx, x.a_data_offset, x.b_data_offset, x.c_data_offset, x.m, x.k, x.n, (double) x.f_ex_sign, (double) x.beta});
    return out;
  }
  void cost(const TenBase *b, int n, const int64_t *aa, const int64_t *ba, double *o) const override {
    auto c = EstimateContractCost(t, cast(b)->t, MakeAxes(n, aa, ba));
    o[0] = c.flops; o[1] = (double) c.gemm_count; o[2] = (double) c.candidate_block_pair_count; o[3] = (double) c.output_block_count;
    o[4] = (double) c.output_raw_elem_count; o[5] = (double) c.read_bytes; o[6] = (double) c.write_bytes; o[7] = (double) c.temp_peak_bytes;
  }
};

template<typename QNT>
IdxBase *MakeIndex(int kind, int dir, int nsct, const int64_t *qnvals, const int64_t *degs) {
  QNSectorVec<QNT> scts;
  for (int s = 0; s < nsct; ++s) scts.push_back(QNSector<QNT>(QNMake<QNT>::make(qnvals + s * QNMake<QNT>::nv), (size_t) degs[s]));
  auto *b = new IdxBox<QNT>(); b->kind = kind;
  b->idx = Index<QNT>(scts, dir < 0 ? TenIndexDirType::IN : TenIndexDirType::OUT);
  return b;
}

template<typename ElemT, typename QNT>
TenBase *MakeTensor(int kind, int dtype, int n, IdxBase *const *idx) {
  IndexVec<QNT> v;
  for (int i = 0; i < n; ++i) v.push_back(static_cast<IdxBox<QNT> *>(idx[i])->idx);
  auto *b = new TenBox<ElemT, QNT>(); b->kind = kind; b->dtype = dtype;
  b->t = QLTensor<ElemT, QNT>(v);
  return b;
}

#define KIND_SWITCH(kind, F, ...)                         \
  switch (kind) {                                         \
    case K_U1: return F<U1QN>(__VA_ARGS__);               \
    case K_FU1: return F<fU1QN>(__VA_ARGS__);             \
    case K_U1U1: return F<U1U1QN>(__VA_ARGS__);           \
    case K_FU1U1: return F<fU1U1QN>(__VA_ARGS__);         \
    case K_Z2: return F<Z2QN>(__VA_ARGS__);               \
    case K_FZ2: return F<fZ2QN>(__VA_ARGS__);             \
    default: return nullptr;                              \
  }

template<typename QNT> TenBase *MakeTensorD(int kind, int dtype, int n, IdxBase *const *idx) {
  if (dtype == D_F64) return MakeTensor<QLTEN_Double, QNT>(kind, dtype, n, idx);
  return MakeTensor<QLTEN_Complex, QNT>(kind, dtype, n, idx);
}

}  // namespace qlref

#endif
