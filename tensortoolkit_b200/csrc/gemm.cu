// gemm.cu -- grouped (ragged-batch) FP64 / complex-FP64 GEMM: all matched sector pairs of one
// contraction in one persistent launch.
//
// Replaces the per-task hp_numeric::MatMultiply calls of the reference's contraction loop
// (include/qlten/qltensor/blk_spar_data_ten/global_operations.h:971-978 ->
//  raw_data_operations.h:530-548 -> framework/hp_numeric/blas_level3.h:35-108 CBLAS / :798-981 cuBLAS):
//     C[m x n] = sign * A[m x k] * B[k x n] + beta * C,   all row-major, lda = k, ldb = n, ldc = n.
// Here the tasks that write the same output block form one group; a CTA owns one BM x BN tile of
// a group and walks ALL of the group's input pairs, accumulating in registers, so C is written
// exactly once, never read, and the summation order is fixed (deterministic, no atomics).
//
// Math: warp-level FP64 tensor-core MMA (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4); tcgen05 has no
// FP64 kind on sm_100a.  Complex pairs use 4 real MMAs on interleaved (re, im) operands.
// Feeding: cp.async (LDGSTS) multi-stage pipeline global -> shared, padded rows so every fragment
// load is bank-conflict free.  Ragged edges are zero-filled by the copy itself (src-size 0).
//
// Narrow pairs (n <= 8 and k <= 32, the MPO-application steps of a DMRG H_eff apply) are
// bandwidth-bound; they go to a one-thread-per-row kernel instead of wasting 128-wide MMA tiles.
#include "common.cuh"

namespace qlb200 {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void CpAsync8(void *smem, const void *gmem, bool pred) {
  const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  const int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void CpAsync16(void *smem, const void *gmem, bool pred) {
  const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void CpAsyncCommit() { asm volatile("cp.async.commit_group;\n" ::); }
template<int N> __device__ __forceinline__ void CpAsyncWait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void Dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// Walks the concatenated k-space of a group's tasks in BK-sized chunks.  The cp.async kernels only
// read row-major operands; a task's block lives either in the caller's buffer or in the workspace.
template<typename T>
struct ChunkCursor {
  uint32_t task, task_end, k0, k;
  const T *a, *b;
  int sign;
  __device__ __forceinline__ void Load(const GemmParams &p) {
    if (task < task_end) {
      const GemmTask t = p.tasks[task];
      a = static_cast<const T *>((t.flags & kTaskASrc) ? p.a_src : p.a_ws) + t.a_off;
      b = static_cast<const T *>((t.flags & kTaskBSrc) ? p.b_src : p.b_ws) + t.b_off;
      k = t.k; sign = t.sign;
    }
  }
  __device__ __forceinline__ bool Valid() const { return task < task_end; }
  template<int BK> __device__ __forceinline__ void Advance(const GemmParams &p) {
    k0 += BK;
    if (k0 >= k) { ++task; k0 = 0; Load(p); }
  }
};

// ================================================================================================
// Real double: CTA tile 128 x 128 x 16, 8 warps as 2 (m) x 4 (n), warp tile 64 x 32.
// ================================================================================================
constexpr int RBM = kRealBM, RBN = kRealBN, RBK = kRealBK, RSTAGES = 4;
constexpr int RLDA = RBK + 4;     // 20 doubles: (row*20 + k) mod 16 distinct over a half-warp
constexpr int RLDB = RBN + 4;     // 132 doubles
constexpr int RA_ELEMS = RBM * RLDA, RB_ELEMS = RBK * RLDB;
constexpr size_t kRealSmem = size_t(RSTAGES) * (RA_ELEMS + RB_ELEMS) * sizeof(double);

__global__ void __launch_bounds__(kThreads, 1)
GemmDmmaReal(GemmParams p) {
  double *__restrict__ C = static_cast<double *>(p.c_out[0]);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sA = reinterpret_cast<double *>(smem_raw);
  double *sB = sA + RSTAGES * RA_ELEMS;
  __shared__ int s_sign[RSTAGES];
  __shared__ uint32_t s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp >> 2) * 64, wn0 = (warp & 3) * 32;
  const int g4 = lane >> 2, t4 = lane & 3;
  // loader coordinates
  const int a_col = tid & 15, a_row = tid >> 4;     // 16 rows per pass, 8 passes
  const int b_col = tid & 127, b_row = tid >> 7;    // 2 rows per pass, 8 passes

  for (;;) {
    if (tid == 0) s_tile = atomicAdd(&p.counters[0], 1u);
    __syncthreads();
    const uint32_t tile_id = s_tile;
    if (tile_id >= p.ntiles) break;
    const GemmTile tile = p.tiles[tile_id];
    const GemmGroup g = p.groups[tile.group];
    const uint32_t row0 = g.row_begin + uint32_t(tile.tm) * RBM, col0 = uint32_t(tile.tn) * RBN;
    const uint32_t m_end = g.row_end, n = g.n;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    ChunkCursor<double> pc;   // producer cursor
    pc.task = g.task_begin; pc.task_end = g.task_end; pc.k0 = 0; pc.k = 0; pc.a = pc.b = nullptr; pc.sign = 1;
    pc.Load(p);
    uint32_t nchunks = 0;
    for (uint32_t t = g.task_begin; t < g.task_end; ++t) nchunks += (p.tasks[t].k + RBK - 1) / RBK;

    auto issue = [&](int stage) {
      if (pc.Valid()) {
        double *dA = sA + stage * RA_ELEMS;
        double *dB = sB + stage * RB_ELEMS;
        const double *gA = pc.a;
        const double *gB = pc.b;
        const uint32_t kk = pc.k0 + a_col;
        const bool kok = kk < pc.k;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const uint32_t row = row0 + a_row + r * 16;
          const bool ok = kok && row < m_end;
          CpAsync8(dA + (a_row + r * 16) * RLDA + a_col, ok ? gA + (unsigned long long) row * pc.k + kk : gA, ok);
        }
        const uint32_t col = col0 + b_col;
        const bool cok = col < n;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const uint32_t krow = pc.k0 + b_row + r * 2;
          const bool ok = cok && krow < pc.k;
          CpAsync8(dB + (b_row + r * 2) * RLDB + b_col, ok ? gB + (unsigned long long) krow * n + col : gB, ok);
        }
        if (tid == 0) s_sign[stage] = pc.sign;
        pc.Advance<RBK>(p);
      }
      CpAsyncCommit();
    };

#pragma unroll
    for (int s = 0; s < RSTAGES - 1; ++s) issue(s);

    for (uint32_t c = 0; c < nchunks; ++c) {
      CpAsyncWait<RSTAGES - 2>();
      __syncthreads();
      issue((c + RSTAGES - 1) % RSTAGES);
      const int stage = c % RSTAGES;
      const double *cA = sA + stage * RA_ELEMS + (wm0 + g4) * RLDA + t4;
      const double *cB = sB + stage * RB_ELEMS + t4 * RLDB + wn0 + g4;
      const bool neg = s_sign[stage] < 0;
#pragma unroll
      for (int ks = 0; ks < RBK / 4; ++ks) {
        double a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] = cA[i * 8 * RLDA + ks * 4]; if (neg) a[i] = -a[i]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = cB[ks * 4 * RLDB + j * 8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) Dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
    CpAsyncWait<0>();

    // epilogue: thread holds C[row = 8i + g4][col = 8j + 2*t4 + {0,1}] of its warp tile
    double *Cg = C + g.c_off;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t row = row0 + wm0 + i * 8 + g4;
      if (row >= m_end) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t col = col0 + wn0 + j * 8 + 2 * t4;
        double *dst = Cg + (unsigned long long) row * n + col;
        if (col < n) dst[0] = acc[i][j][0];
        if (col + 1 < n) dst[1] = acc[i][j][1];
      }
    }
    __syncthreads();   // all smem reads of this tile done before the next tile's prologue overwrites
  }
  // self-resetting tile counter: the last CTA to leave zeroes both counters for the next launch
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&p.counters[1], 1u) == gridDim.x - 1) { p.counters[0] = 0; p.counters[1] = 0; __threadfence(); }
  }
}

// ================================================================================================
// Complex double (interleaved re, im): CTA tile 64 x 128 x 8, 8 warps as 2 x 4, warp tile 32 x 32.
// ================================================================================================
constexpr int CBM = kCplxBM, CBN = kCplxBN, CBK = kCplxBK, CSTAGES = 4;
constexpr int CLDA = CBK + 4;     // 12 complex: (row*12 + k) mod 8 distinct over a quarter-warp
constexpr int CLDB = CBN + 2;     // 130 complex: (k*130 + n) mod 8 = 2k + n distinct
constexpr int CA_ELEMS = CBM * CLDA, CB_ELEMS = CBK * CLDB;
constexpr size_t kCplxSmem = size_t(CSTAGES) * (CA_ELEMS + CB_ELEMS) * sizeof(double2);

__global__ void __launch_bounds__(kThreads, 1)
GemmDmmaCplx(GemmParams p) {
  double2 *__restrict__ C = static_cast<double2 *>(p.c_out[0]);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *sA = reinterpret_cast<double2 *>(smem_raw);
  double2 *sB = sA + CSTAGES * CA_ELEMS;
  __shared__ int s_sign[CSTAGES];
  __shared__ uint32_t s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;
  const int g4 = lane >> 2, t4 = lane & 3;
  const int a_col = tid & 7, a_row = tid >> 3;      // 32 rows per pass, 2 passes
  const int b_col = tid & 127, b_row = tid >> 7;    // 2 rows per pass, 4 passes

  for (;;) {
    if (tid == 0) s_tile = atomicAdd(&p.counters[0], 1u);
    __syncthreads();
    const uint32_t tile_id = s_tile;
    if (tile_id >= p.ntiles) break;
    const GemmTile tile = p.tiles[tile_id];
    const GemmGroup g = p.groups[tile.group];
    const uint32_t row0 = g.row_begin + uint32_t(tile.tm) * CBM, col0 = uint32_t(tile.tn) * CBN;
    const uint32_t m_end = g.row_end, n = g.n;

    double cr[4][4][2], ci[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;

    ChunkCursor<double2> pc;
    pc.task = g.task_begin; pc.task_end = g.task_end; pc.k0 = 0; pc.k = 0; pc.a = pc.b = nullptr; pc.sign = 1;
    pc.Load(p);
    uint32_t nchunks = 0;
    for (uint32_t t = g.task_begin; t < g.task_end; ++t) nchunks += (p.tasks[t].k + CBK - 1) / CBK;

    auto issue = [&](int stage) {
      if (pc.Valid()) {
        double2 *dA = sA + stage * CA_ELEMS;
        double2 *dB = sB + stage * CB_ELEMS;
        const double2 *gA = pc.a;
        const double2 *gB = pc.b;
        const uint32_t kk = pc.k0 + a_col;
        const bool kok = kk < pc.k;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t row = row0 + a_row + r * 32;
          const bool ok = kok && row < m_end;
          CpAsync16(dA + (a_row + r * 32) * CLDA + a_col, ok ? gA + (unsigned long long) row * pc.k + kk : gA, ok);
        }
        const uint32_t col = col0 + b_col;
        const bool cok = col < n;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const uint32_t krow = pc.k0 + b_row + r * 2;
          const bool ok = cok && krow < pc.k;
          CpAsync16(dB + (b_row + r * 2) * CLDB + b_col, ok ? gB + (unsigned long long) krow * n + col : gB, ok);
        }
        if (tid == 0) s_sign[stage] = pc.sign;
        pc.Advance<CBK>(p);
      }
      CpAsyncCommit();
    };

#pragma unroll
    for (int s = 0; s < CSTAGES - 1; ++s) issue(s);

    for (uint32_t c = 0; c < nchunks; ++c) {
      CpAsyncWait<CSTAGES - 2>();
      __syncthreads();
      issue((c + CSTAGES - 1) % CSTAGES);
      const int stage = c % CSTAGES;
      const double2 *cA = sA + stage * CA_ELEMS + (wm0 + g4) * CLDA + t4;
      const double2 *cB = sB + stage * CB_ELEMS + t4 * CLDB + wn0 + g4;
      const bool neg = s_sign[stage] < 0;
#pragma unroll
      for (int ks = 0; ks < CBK / 4; ++ks) {
        double2 a[4], b[4];
        double nai[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a[i] = cA[i * 8 * CLDA + ks * 4];
          if (neg) { a[i].x = -a[i].x; a[i].y = -a[i].y; }
          nai[i] = -a[i].y;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = cB[ks * 4 * CLDB + j * 8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            Dmma(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
            Dmma(ci[i][j][0], ci[i][j][1], a[i].x, b[j].y);
          }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            Dmma(cr[i][j][0], cr[i][j][1], nai[i], b[j].y);
            Dmma(ci[i][j][0], ci[i][j][1], a[i].y, b[j].x);
          }
      }
    }
    CpAsyncWait<0>();

    double2 *Cg = C + g.c_off;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t row = row0 + wm0 + i * 8 + g4;
      if (row >= m_end) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t col = col0 + wn0 + j * 8 + 2 * t4;
        double2 *dst = Cg + (unsigned long long) row * n + col;
        if (col < n) dst[0] = make_double2(cr[i][j][0], ci[i][j][0]);
        if (col + 1 < n) dst[1] = make_double2(cr[i][j][1], ci[i][j][1]);
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&p.counters[1], 1u) == gridDim.x - 1) { p.counters[0] = 0; p.counters[1] = 0; __threadfence(); }
  }
}

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
__global__ void __launch_bounds__(kSkinnyThreads, kSkinnyMinCtas)
GemmSkinny(GemmParams p) {
  using E = Elem<CPLX>;
  using T = typename E::T;
  __shared__ const T *s_ap[kSkinnyTermChunk];
  __shared__ uint32_t s_as[kSkinnyTermChunk];
  __shared__ T s_coef[kSkinnyTermChunk][kSkinnyMaxN];
  __shared__ uint32_t s_nterm;
  const uint32_t tid = threadIdx.x;
  for (uint32_t it = blockIdx.x; it < p.nitems; it += gridDim.x) {
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
#pragma unroll 1
        for (uint32_t x = 0; x < nterm; ++x) {
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
  cudaError_t e = cudaFuncSetAttribute(GemmDmmaReal, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kRealSmem));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(GemmDmmaCplx, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kCplxSmem));
  if (e != cudaSuccess) return e;
  e = ConfigureWsKernel();
  if (e != cudaSuccess) return e;
  return ConfigureWsRealKernel();
}

cudaError_t LaunchGemmDmma(int dtype, const GemmParams &p, int num_sms, cudaStream_t stream) {
  if (p.ntiles == 0) return cudaSuccess;
  const uint32_t grid = p.ntiles < uint32_t(num_sms) ? p.ntiles : uint32_t(num_sms);
  if (dtype == 0) GemmDmmaReal<<<grid, kThreads, kRealSmem, stream>>>(p);
  else GemmDmmaCplx<<<grid, kThreads, kCplxSmem, stream>>>(p);
  return cudaGetLastError();
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
