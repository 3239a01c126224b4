// gemm.cu -- the narrow-pair kernel of the grouped GEMM and the kernel configuration entry point.
//
// The grouped (ragged-batch) GEMM replaces the per-task hp_numeric::MatMultiply calls of the reference's contraction loop
// (include/qlten/qltensor/blk_spar_data_ten/global_operations.h:971-978 ->
//  raw_data_operations.h:530-548 -> framework/hp_numeric/blas_level3.h:35-108 CBLAS / :798-981 cuBLAS):
//     C[m x n] = sign * A[m x k] * B[k x n] + beta * C,   all row-major, lda = k, ldb = n, ldc = n.
// The tasks that write the same output block form one group; C is written exactly once, never read, and the summation
// order is fixed (deterministic, no atomics).  GEMM-shaped groups run on the FP64 tensor pipe in the warp-specialised
// kernels (gemm_ws.cu complex, gemm_ws_real.cu double); narrow pairs (n <= 8 and k <= 32, the MPO-application steps of a
// DMRG H_eff apply) are bandwidth-bound and run here.
#include "common.cuh"

namespace qlb200 {

namespace {


// ================================================================================================
// Narrow pairs (n <= kSkinnyMaxN, k <= kSkinnyMaxK): HBM-bound, reads m*k per pair, writes m*n once.
//
// A work item is a run of rows of one output block holding about kSkinnyElems output elements; one
// thread owns up to kSkinnyPerThread of them (element e = tid + 256*u: consecutive threads write
// consecutive elements of C).  The (pair, kk) terms of the block are first flattened into a small
// shared-memory table {A pointer, A row stride, B coefficients with the sign folded in}; the main
// loop then issues one A load per owned element per term -- kSkinnyPerThread independent loads in flight per
// thread -- and multiplies by coefficients broadcast from shared memory.  A and B blocks are read in
// place: row-major, or 2-D transposed in the caller's buffer (kTask?Trans).
// ================================================================================================
template<bool CPLX> struct Elem;
template<> struct Elem<false> {
  using T = double;
  static __device__ __forceinline__ T Zero() { return 0.0; }
  static __device__ __forceinline__ void Fma(T &acc, T a, T b) { acc = fma(a, b, acc); }
  static __device__ __forceinline__ T Signed(T v, int sign) { return sign < 0 ? -v : v; }
};
template<> struct Elem<true> {
  using T = double2;
  static __device__ __forceinline__ T Zero() { return make_double2(0.0, 0.0); }
  static __device__ __forceinline__ void Fma(T &acc, T a, T b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
  }
  static __device__ __forceinline__ T Signed(T v, int sign) { return sign < 0 ? make_double2(-v.x, -v.y) : v; }
};

constexpr int kSkinnyTermChunk = 64;

template<bool CPLX, bool ACC>
__global__ void __launch_bounds__(kSkinnyThreads, CPLX ? kSkinnyMinCtas - 1 : kSkinnyMinCtas)     // complex: 8 double2 loads in flight need > 64 registers
GemmSkinny(GemmParams p) {
  using E = Elem<CPLX>;
  using T = typename E::T;
  __shared__ const T *s_ap[kSkinnyTermChunk];
  __shared__ uint32_t s_as[kSkinnyTermChunk];
  __shared__ T s_coef[kSkinnyTermChunk][kSkinnyMaxN];
  __shared__ uint32_t s_nterm;
  const uint32_t tid = threadIdx.x;
  for (uint32_t it = blockIdx.x; it < p.nitems; it += gridDim.x) {
    // (prefetching the next item's descriptors under this item's main loop was measured and changes nothing: with 3-4
    // resident CTAs per SM the chain of one CTA hides behind the streaming of the others -- exp/r2_call21.sh)
    const SkinnyItem item = p.items[it];
    const GemmGroup g = p.groups[item.group];
    const uint32_t n = g.n;
    // An item holds p.skinny_sub sub-chunks of about kSkinnyElems outputs: fetching the item / group / task descriptors is a
    // chain of dependent global loads (microseconds), so it is paid once per item, and when all (pair, kk) terms of the block
    // fit one table chunk -- every MPO-step block does -- the table is built once and reused by all sub-chunks.
    const uint32_t sub_rows = uint32_t(kSkinnyElems) / n;
    const uint32_t item_rows = min(sub_rows * p.skinny_sub, g.row_end - item.row0);
    bool table_valid = false;      // the shared table holds ALL terms of this block
    for (uint32_t sr = 0; sr < item_rows; sr += sub_rows) {
      const uint32_t row_base = item.row0 + sr;
      const uint32_t rows = min(sub_rows, item_rows - sr);
      const uint32_t total = rows * n;
      uint32_t row[kSkinnyPerThread], col[kSkinnyPerThread];
      T acc[kSkinnyPerThread];
#pragma unroll
      for (int u = 0; u < kSkinnyPerThread; ++u) {
        const uint32_t e = tid + u * kSkinnyThreads;
        const uint32_t r = e / n;
        row[u] = row_base + r; col[u] = e - r * n;
        acc[u] = E::Zero();
      }
      // walk the group's (pair, kk) terms in chunks of kSkinnyTermChunk
      uint32_t t = g.task_begin, kk0 = 0;
      uint32_t nterm_kept = 0;
      while (t < g.task_end) {
        uint32_t nterm = 0;
        if (!table_valid) {
          __syncthreads();   // previous chunk (or previous item) fully consumed
          // every thread walks the same short task list; thread x fills term x
          uint32_t tt = t, kk = kk0;
          while (tt < g.task_end && nterm < uint32_t(kSkinnyTermChunk)) {
            const GemmTask tk = p.tasks[tt];
            const uint32_t take = min(tk.k - kk, uint32_t(kSkinnyTermChunk) - nterm);
            if (tid >= nterm && tid < nterm + take) {
              const uint32_t k1 = kk + (tid - nterm);
              const T *a = static_cast<const T *>((tk.flags & kTaskASrc) ? p.a_src : p.a_ws) + tk.a_off;
              const T *b = static_cast<const T *>((tk.flags & kTaskBSrc) ? p.b_src : p.b_ws) + tk.b_off;
              if (tk.flags & kTaskATrans) { s_ap[tid] = a + (unsigned long long) k1 * g.m; s_as[tid] = 1; }
              else { s_ap[tid] = a + k1; s_as[tid] = tk.k; }
              for (uint32_t j = 0; j < n; ++j) {
                // B(k1, j): stored n x k (transposed), row-major k x n (b_run >= n: no division), or a strided view of the block
                const T v = (tk.flags & kTaskBTrans) ? b[(unsigned long long) j * tk.k + k1]
                            : tk.b_run >= n         ? b[(unsigned long long) k1 * tk.b_rs + j]
                                                     : b[(unsigned long long) k1 * tk.b_rs + (j / tk.b_run) * tk.b_cs + j % tk.b_run];
                s_coef[tid][j] = E::Signed(v, tk.sign);
              }
            }
            nterm += take; kk += take;
            if (kk >= tk.k) { ++tt; kk = 0; }
          }
          // first chunk reached the end of the task list: the table is complete and serves every sub-chunk of the item
          if (t == g.task_begin && kk0 == 0 && tt >= g.task_end) { table_valid = true; s_nterm = nterm; }
          t = tt; kk0 = kk;
          __syncthreads();
        } else {
          nterm = nterm_kept = s_nterm;
          t = g.task_end;
        }
        // two terms per trip: 2 x kSkinnyPerThread independent loads in flight per thread
        uint32_t x = 0;
#pragma unroll 1
        for (; x + 2 <= nterm; x += 2) {
          const T *ap0 = s_ap[x], *ap1 = s_ap[x + 1];
          const unsigned long long as0 = s_as[x], as1 = s_as[x + 1];
          T av0[kSkinnyPerThread], av1[kSkinnyPerThread];
#pragma unroll
          for (int u = 0; u < kSkinnyPerThread; ++u)
            if (tid + u * kSkinnyThreads < total) { av0[u] = ap0[row[u] * as0]; av1[u] = ap1[row[u] * as1]; }
#pragma unroll
          for (int u = 0; u < kSkinnyPerThread; ++u)
            if (tid + u * kSkinnyThreads < total) { E::Fma(acc[u], av0[u], s_coef[x][col[u]]); E::Fma(acc[u], av1[u], s_coef[x + 1][col[u]]); }
        }
        if (x < nterm) {
          const T *ap = s_ap[x];
          const unsigned long long as = s_as[x];
          T av[kSkinnyPerThread];
#pragma unroll
          for (int u = 0; u < kSkinnyPerThread; ++u)
            if (tid + u * kSkinnyThreads < total) av[u] = ap[row[u] * as];
#pragma unroll
          for (int u = 0; u < kSkinnyPerThread; ++u)
            if (tid + u * kSkinnyThreads < total) E::Fma(acc[u], av[u], s_coef[x][col[u]]);
        }
      }
      (void) nterm_kept;
      const unsigned long long cbase = g.c_off + (unsigned long long) row_base * n;
      for (uint32_t d = 0; d < p.n_out; ++d) {
        T *cb = static_cast<T *>(p.c_out[d]) + cbase;
        const T *ci = static_cast<const T *>(p.c_in) + g.c_in_off + (unsigned long long) row_base * n;
#pragma unroll
        for (int u = 0; u < kSkinnyPerThread; ++u) {
          const uint32_t e = tid + u * kSkinnyThreads;
          if constexpr (ACC) {
            if (e < total) StoreOut(cb + e, AxpbyOut(p, acc[u], ci + e, g.beta_on != 0), p.mcast);
          } else {
            if (e < total) StoreOut(cb + e, acc[u], p.mcast);
          }
        }
      }
    }
  }
}

}  // namespace

cudaError_t ConfigureKernels() {
  cudaError_t e = ConfigureWsKernel();
  if (e != cudaSuccess) return e;
  return ConfigureWsRealKernel();
}

cudaError_t LaunchGemmSkinny(int dtype, const GemmParams &p, int num_sms, cudaStream_t stream) {
  if (p.nitems == 0) return cudaSuccess;
  const uint32_t cap = uint32_t(num_sms) * 8u;
  const uint32_t grid = p.nitems < cap ? p.nitems : cap;
  if (p.accum) {
    if (dtype == 0) GemmSkinny<false, true><<<grid, kSkinnyThreads, 0, stream>>>(p);
    else GemmSkinny<true, true><<<grid, kSkinnyThreads, 0, stream>>>(p);
  } else {
    if (dtype == 0) GemmSkinny<false, false><<<grid, kSkinnyThreads, 0, stream>>>(p);
    else GemmSkinny<true, false><<<grid, kSkinnyThreads, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace qlb200
