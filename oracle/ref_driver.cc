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

int qlref_tensor_write(const void *t, const char *path) { return static_cast<const TenBase *>(t)->write_file(path); }
int qlref_tensor_read(void *t, const char *path) { return static_cast<TenBase *>(t)->read_file(path); }
void *qlref_contract_contiguous(const void *a, const void *b, int64_t a_start, int64_t b_start, int64_t size, int side) {
  return static_cast<const TenBase *>(a)->contract_contiguous(static_cast<const TenBase *>(b), a_start, b_start, size, side);
}

}  // extern "C"
