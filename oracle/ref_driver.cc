// oracle/ref_driver.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Thin C ABI over the UNMODIFIED reference (TensorToolkit headers under /root/reference/include,
// vendored HPTT, OpenBLAS) so that tests and the cpu_baseline leg of bench.py can build reference
// tensors, run the reference's own qlten::Contract / Transpose on the host and read back block
// structure, task tables and raw data.  Built by oracle/Makefile into oracle/_ref/libqlref.so.
//
// This library contains the reference ONLY (it links nothing of the product, so the reference arm of bench.py maps no
// product code); the drop-in adapter instantiated on the same tensor types lives in adapter_driver.cc / libqladapter.so.
#include "ref_handles.h"

using namespace qlref;

namespace {

// ---- tensors built from the product's host mirror (same blocks, same raw data) ----------------
// Inserts the listed blocks (DataBlksInsert, allocating) and copies `raw` (blocks in ascending blk_idx order, the
// reference's own layout) into the tensor.  Returns 0, or -1 when the sizes disagree.
template<typename Box>
int FillBlocks(Box *bx, uint64_t nblk, const uint32_t *coors, const void *raw, uint64_t nelem) {
  auto &t = bx->t;
  const size_t r = t.Rank();
  auto &bsdt = t.GetBlkSparDataTen();
  if (r == 0) {                                   // scalar: ElemSet creates the size-1 raw buffer
    if (nelem != 1) return -1;
    t.SetElem({}, *static_cast<const typename Box::Elem *>(raw));
    return 0;
  }
  std::vector<CoorsT> cs(nblk, CoorsT(r));
  std::vector<size_t> idxs(nblk);
  for (uint64_t b = 0; b < nblk; ++b) {
    for (size_t i = 0; i < r; ++i) cs[b][i] = coors[b * r + i];
    idxs[b] = bsdt.BlkCoorsToBlkIdx(cs[b]);
  }
  bsdt.DataBlksInsert(idxs, cs, true, false);
  if (bsdt.GetActualRawDataSize() != nelem) return -1;
  std::memcpy(bsdt.GetActualRawDataPtr(), raw, nelem * sizeof(typename Box::Elem));
  return 0;
}

// ---- the executor loop on bare descriptor tables (BASELINE configs[4], the ragged stress test) --
// Exactly the reference's CtrctTwoBSDTAndAssignIn (global_operations.h:895-992) without the QLTensor shells: tasks
// sorted by (c, beta); every distinct A / B block transposed once with hp_numeric::TensorTranspose (HPTT) into its own
// buffer; one hp_numeric::MatMultiply (CBLAS) per task with alpha = sign and beta = 0 for the first task of an output
// block, 1 for the rest.  tu = [ntask][8] {a_ord, b_ord, a_off, b_off, c_off, m, k, n}; ts = [ntask][2] {sign, beta}.
// If a_t / b_t are given, the transposed copy of block b is also stored at a_t + a_off[b] (bit-exact permute oracle).
template<typename ElemT>
void RawContract(int a_rank, const int32_t *a_perm, const uint32_t *a_shape, const uint64_t *a_off, int b_rank,
                        const int32_t *b_perm, const uint32_t *b_shape, const uint64_t *b_off, uint64_t ntask, const uint64_t *tu,
                        const double *ts, const ElemT *A, const ElemT *B, ElemT *C, ElemT *a_t, ElemT *b_t) {
  std::vector<uint64_t> order(ntask);
  for (uint64_t i = 0; i < ntask; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](uint64_t x, uint64_t y) {
    if (tu[8 * x + 4] != tu[8 * y + 4]) return tu[8 * x + 4] < tu[8 * y + 4];
    return ts[2 * x + 1] < ts[2 * y + 1];
  });
  auto transposed = [](int rank, const int32_t *perm, const uint32_t *shape, const ElemT *src, ElemT *keep) {
    std::vector<int> ord(perm, perm + rank);
    std::vector<size_t> shp(shape, shape + rank);
    std::vector<int> tshape(rank);
    size_t size = 1;
    for (int i = 0; i < rank; ++i) { tshape[i] = (int) shape[perm[i]]; size *= shape[i]; }
    ElemT *out = (ElemT *) qlten::QLMalloc(size * sizeof(ElemT));
    hp_numeric::TensorTranspose(ord, (size_t) rank, const_cast<ElemT *>(src), shp, out, tshape, 1.0);
    if (keep) std::memcpy(keep, out, size * sizeof(ElemT));
    return out;
  };
  std::unordered_map<uint64_t, ElemT *> ta, tb;
  for (uint64_t oi = 0; oi < ntask; ++oi) {
    const uint64_t *u = tu + 8 * order[oi];
    const double sign = ts[2 * order[oi]], beta = ts[2 * order[oi] + 1];
    const ElemT *a = A + u[2], *b = B + u[3];
    if (a_perm) {
      auto it = ta.find(u[0]);
      if (it == ta.end()) it = ta.emplace(u[0], transposed(a_rank, a_perm, a_shape + u[0] * a_rank, a, a_t ? a_t + a_off[u[0]] : nullptr)).first;
      a = it->second;
    }
    if (b_perm) {
      auto it = tb.find(u[1]);
      if (it == tb.end()) it = tb.emplace(u[1], transposed(b_rank, b_perm, b_shape + u[1] * b_rank, b, b_t ? b_t + b_off[u[1]] : nullptr)).first;
      b = it->second;
    }
    hp_numeric::MatMultiply(sign, a, b, (size_t) u[5], (size_t) u[6], (size_t) u[7], ElemT(beta), C + u[4]);
  }
  for (auto &kv : ta) qlten::QLFree(kv.second);
  for (auto &kv : tb) qlten::QLFree(kv.second);
}

}  // namespace

extern "C" {

void qlref_set_seed(uint64_t seed) { SetRandomSeed((size_t) seed); }
void qlref_set_threads(int n) { hp_numeric::SetTensorManipulationThreads((unsigned) n); }

void *qlref_index_new(int kind, int dir, int nsct, const int64_t *qnvals, const int64_t *degs) {
  KIND_SWITCH(kind, MakeIndex, kind, dir, nsct, qnvals, degs)
}
void qlref_index_free(void *p) { delete static_cast<IdxBase *>(p); }

void *qlref_tensor_new(int kind, int dtype, int nidx, void *const *idx) {
  KIND_SWITCH(kind, MakeTensorD, kind, dtype, nidx, reinterpret_cast<IdxBase *const *>(idx))
}
void qlref_tensor_free(void *t) { delete static_cast<TenBase *>(t); }
void *qlref_tensor_clone(const void *t) { return static_cast<const TenBase *>(t)->clone(); }
int qlref_tensor_rank(const void *t) { return static_cast<const TenBase *>(t)->rank(); }
int qlref_tensor_dtype(const void *t) { return static_cast<const TenBase *>(t)->dtype; }
int qlref_tensor_is_default(const void *t) { return static_cast<const TenBase *>(t)->is_default(); }
int qlref_tensor_fermionic(const void *t) { return static_cast<const TenBase *>(t)->fermionic(); }
uint64_t qlref_tensor_nblk(const void *t) { return static_cast<const TenBase *>(t)->nblk(); }
uint64_t qlref_tensor_raw_size(const void *t) { return static_cast<const TenBase *>(t)->raw_size(); }
const void *qlref_tensor_raw(const void *t) { return static_cast<const TenBase *>(t)->raw(); }
void qlref_tensor_shell(const void *t, uint32_t *nsct, int8_t *dirs) {
  static_cast<const TenBase *>(t)->nsct(nsct); static_cast<const TenBase *>(t)->dirs(dirs);
}
void qlref_tensor_sectors(const void *t, uint32_t *degs, uint8_t *parities) {
  static_cast<const TenBase *>(t)->degs(degs); static_cast<const TenBase *>(t)->parities(parities);
}
void qlref_tensor_blocks(const void *t, uint64_t *idx, uint32_t *coors, uint32_t *shape, uint64_t *off) {
  static_cast<const TenBase *>(t)->blocks(idx, coors, shape, off);
}
void qlref_tensor_random(void *t, const int64_t *div) { static_cast<TenBase *>(t)->random(div); }
void qlref_tensor_transpose(void *t, const int64_t *perm) { static_cast<TenBase *>(t)->transpose(perm); }
double qlref_tensor_norm2(const void *t) { return static_cast<const TenBase *>(t)->norm2(); }
int qlref_tensor_indexes_equal(const void *a, const void *b) { return static_cast<const TenBase *>(a)->indexes_equal(static_cast<const TenBase *>(b)); }

void *qlref_contract(const void *a, const void *b, int n, const int64_t *aa, const int64_t *ba) {
  return static_cast<const TenBase *>(a)->contract(static_cast<const TenBase *>(b), n, aa, ba);
}
// wall seconds of `reps` reference Contract calls (inputs pre-built, result discarded), best of reps
double qlref_contract_time(const void *a, const void *b, int n, const int64_t *aa, const int64_t *ba, int reps) {
  double best = 1e300;
  for (int r = 0; r < reps; ++r) {
    auto t0 = std::chrono::steady_clock::now();
    TenBase *c = static_cast<const TenBase *>(a)->contract(static_cast<const TenBase *>(b), n, aa, ba);
    auto t1 = std::chrono::steady_clock::now();
    delete c;
    best = std::min(best, std::chrono::duration<double>(t1 - t0).count());
  }
  return best;
}
void *qlref_contract_1sector(const void *a, int64_t axis, int64_t sct, const void *b, int n, const int64_t *aa, const int64_t *ba) {
  return static_cast<const TenBase *>(a)->contract_1sector(axis, sct, static_cast<const TenBase *>(b), n, aa, ba);
}
uint64_t qlref_contract_tasks(const void *a, const void *b, int n, const int64_t *aa, const int64_t *ba, int sorted,
                              uint64_t cap, uint64_t *u9, double *d2) {
  auto v = static_cast<const TenBase *>(a)->tasks(static_cast<const TenBase *>(b), n, aa, ba, sorted != 0);
  for (uint64_t i = 0; i < v.size() && i < cap; ++i) {
    const auto &x = v[i];
    uint64_t *u = u9 + 9 * i;
    u[0] = x.a_blk_idx; u[1] = x.b_blk_idx; u[2] = x.c_blk_idx; u[3] = x.a_off; u[4] = x.b_off; u[5] = x.c_off; u[6] = x.m; u[7] = x.k; u[8] = x.n;
    d2[2 * i] = x.sign; d2[2 * i + 1] = x.beta;
  }
  return v.size();
}
void qlref_contract_cost(const void *a, const void *b, int n, const int64_t *aa, const int64_t *ba, double *out8) {
  static_cast<const TenBase *>(a)->cost(static_cast<const TenBase *>(b), n, aa, ba, out8);
}

// a default-constructed tensor of the same element / quantum-number type as `like` (output of the accumulate forms)
void *qlref_tensor_new_default(const void *like) {
  const TenBase *l = static_cast<const TenBase *>(like);
  TenBase *c = l->clone();
  int64_t none = 0;
  (void) none;
  c->reset_default();
  return c;
}
// returns 1 = done, 0 = the Try... probe reported a layout mismatch, -1 = exception (text on stderr)
int qlref_contract_accumulate(const void *a, const void *b, int64_t a_start, int64_t b_start, int64_t size, const double *alpha2,
                              const double *beta2, void *c, int try_only, uint64_t *stats10) {
  try {
    return static_cast<const TenBase *>(a)->contract_accumulate(static_cast<const TenBase *>(b), a_start, b_start, size, alpha2, beta2,
                                                                static_cast<TenBase *>(c), try_only, stats10);
  } catch (const std::exception &e) { std::fprintf(stderr, "qlref_contract_accumulate: %s\n", e.what()); return -1; }
}
void *qlref_apply_rank2(const void *x, const void *op1, int64_t axis1, const void *op2, int64_t axis2) {
  try {
    return static_cast<const TenBase *>(x)->apply_rank2(static_cast<const TenBase *>(op1), axis1, static_cast<const TenBase *>(op2), axis2);
  } catch (const std::exception &e) { std::fprintf(stderr, "qlref_apply_rank2: %s\n", e.what()); return nullptr; }
}
int qlref_tensor_write(const void *t, const char *path) { return static_cast<const TenBase *>(t)->write_file(path); }
int qlref_tensor_read(void *t, const char *path) { return static_cast<TenBase *>(t)->read_file(path); }
void *qlref_contract_contiguous(const void *a, const void *b, int64_t a_start, int64_t b_start, int64_t size, int side) {
  return static_cast<const TenBase *>(a)->contract_contiguous(static_cast<const TenBase *>(b), a_start, b_start, size, side);
}

int qlref_tensor_fill(void *t, uint64_t nblk, const uint32_t *coors, const void *raw, uint64_t nelem) {
  TenBase *tb = static_cast<TenBase *>(t);
#define QLREF_FILL(K, QN)                                                                                           \
  case K:                                                                                                           \
    if (tb->dtype == D_F64) return FillBlocks(static_cast<TenBox<QLTEN_Double, QN> *>(tb), nblk, coors, raw, nelem); \
    return FillBlocks(static_cast<TenBox<QLTEN_Complex, QN> *>(tb), nblk, coors, raw, nelem);
  switch (tb->kind) {
    QLREF_FILL(K_U1, U1QN) QLREF_FILL(K_FU1, fU1QN) QLREF_FILL(K_U1U1, U1U1QN)
    QLREF_FILL(K_FU1U1, fU1U1QN) QLREF_FILL(K_Z2, Z2QN) QLREF_FILL(K_FZ2, fZ2QN)
    default: return -1;
  }
#undef QLREF_FILL
}

double qlref_raw_contract(int dtype, int a_rank, const int32_t *a_perm, const uint32_t *a_shape, const uint64_t *a_off, int b_rank,
                          const int32_t *b_perm, const uint32_t *b_shape, const uint64_t *b_off, uint64_t ntask, const uint64_t *tu,
                          const double *ts, const void *A, const void *B, void *C, void *a_t, void *b_t) {
  auto t0 = std::chrono::steady_clock::now();
  if (dtype == D_F64)
    RawContract<QLTEN_Double>(a_rank, a_perm, a_shape, a_off, b_rank, b_perm, b_shape, b_off, ntask, tu, ts, (const QLTEN_Double *) A,
                              (const QLTEN_Double *) B, (QLTEN_Double *) C, (QLTEN_Double *) a_t, (QLTEN_Double *) b_t);
  else
    RawContract<QLTEN_Complex>(a_rank, a_perm, a_shape, a_off, b_rank, b_perm, b_shape, b_off, ntask, tu, ts, (const QLTEN_Complex *) A,
                               (const QLTEN_Complex *) B, (QLTEN_Complex *) C, (QLTEN_Complex *) a_t, (QLTEN_Complex *) b_t);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
