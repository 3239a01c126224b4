// gemm_ws_real.cu -- warp-specialised grouped FP64 GEMM for real double tensors.
//
// Same pipeline as the complex kernel (gemm_ws.cu: 4 consumer warps + 4 cp.async producer warps,
// full/empty mbarrier ring, two CTAs per SM, dynamic tile counter, operands read in place row-major
// or 2-D transposed, ragged tiles specialised at compile time); the differences are the shapes:
//   * CTA tile 64 x 128, k-stage 16: a warp owns all 64 rows (8 m8 groups) and the n8 column groups
//     {q, q+4, q+8, q+12}: 32 DMMA.8x8x4 per k4 step, 64 accumulator doubles per lane -- the same
//     MMA count per stage and the same register budget as the complex kernel;
//   * elements are 8 bytes, so copies are 8-byte cp.async (any block offset / leading dimension is
//     legal) and a fragment load is an LDS.64 served per half-warp (lanes g4 in {4h..4h+3}, t4 in 0..3):
//       A row-major   [64 m][20]    (20*g4 + t4)  mod 16 = 4*g4 + t4 distinct
//       A transposed  [16 k][68]    (68*t4 + g4)  mod 16 = 4*t4 + g4 distinct
//       B row-major   [16 k][132]   (132*t4 + g4) mod 16 = 4*t4 + g4 distinct
//       B transposed  [128 n][16] with k ^ 4*(n&3): 4*((g4&3) ^ ks) + t4 distinct
#include "common.cuh"
#include "ws_common.cuh"
// L2 prefetch distance of the producer for ROW-MAJOR A operands, in stages (0 = off).  The rows of such a tile lie a whole block
// row (k elements) apart: every stage touches 64 different DRAM pages for 128 bytes each, and it is complete only when its
// SLOWEST line has landed.  One lane per row asks L2 (prefetch.global.L2, SASS CCTL.E.PF2) for the line the row needs kPf
// stages from now; the copy that comes for it later finds a short, uniform latency.  Measured (exp/r2_call23.sh): step 4 of the
// U(1) chain 0.76 -> 0.82 of DGEMM at D=4096, 0.79 -> 0.85 on the Hubbard chain; the distance (4 / 8 / 16) does not matter.
// Prefetching the other three operand layouts as well (k x m stored A, both B layouts: 16 rows of 512 .. 1024 contiguous
// bytes per stage, or an L2-resident operand) was measured too and LOSES 1 - 4 % (exp/r2_call24.sh); k x m stored A alone
// changes nothing (exp/r2_call44.sh): not done.  The complex
// kernel (16-byte elements: 256 bytes per row and stage) gains nothing from it either (exp/r2_call25.sh).
#ifndef QLB200_REAL_PF_STAGES
#define QLB200_REAL_PF_STAGES 8
#endif
#ifdef QLB200_EXP_NOCOPY
#define CpAsync8Z(a, b, c) ((void) 0)
#endif

namespace qlb200 {

namespace {

constexpr int RBM = kWsRealBM, RBN = kWsRealBN, RBK = kWsRealBK;
constexpr int RLDA = RBK + 4, RLDAT = RBM + 4, RLDB = RBN + 4;
constexpr int RA_ELEMS = RBM * RLDA, RB_ELEMS = RBK * RLDB, RSTAGE_ELEMS = RA_ELEMS + RB_ELEMS;
static_assert(RBK * RLDAT <= RA_ELEMS && RBN * RBK <= RB_ELEMS, "transposed tiles must fit the stage");
constexpr int kRealStages = 4;
constexpr size_t kRealWsSmem = size_t(kRealStages) * RSTAGE_ELEMS * sizeof(double) + 2 * kRealStages * sizeof(uint64_t) +
                               kRealStages * sizeof(StageMeta);

struct FragAddrR {
  const double *a, *b;
  uint32_t a_i, a_k;        // A: + i * a_i (m8 group) + ks * a_k
  uint32_t b_j, b_k, b_x;   // B: + j * b_j (owned n8 group) + (ks ^ b_x) * b_k
};

// One k4 step of a warp's sub-tile (MT valid m8 row groups x NT valid n8 column groups), fragments at pa + i * a_i and
// pb + j * b_j.  The pair's sign is not applied here: the accumulators live in the sign frame of the current pair and
// are flipped once where the frame changes (kFlagFlip, see gemm_ws.cu).
template<int MT, int NT>
__device__ __forceinline__ void KStepR(double (&acc)[8][4][2], const double *pa, uint32_t a_i, const double *pb, uint32_t b_j) {
  double a[MT], b[NT];
#pragma unroll
#ifdef QLB200_EXP_NOLDS
  for (int i = 0; i < MT; ++i) a[i] = __longlong_as_double(0x3ff0000000000000ll + (long long) (size_t) pa + i);
#pragma unroll
  for (int j = 0; j < NT; ++j) b[j] = __longlong_as_double(0x3ff0000000000000ll + (long long) (size_t) pb + j);
#else
  for (int i = 0; i < MT; ++i) a[i] = pa[i * a_i];
#pragma unroll
  for (int j = 0; j < NT; ++j) b[j] = pb[j * b_j];
#endif
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) DmmaNv(acc[i][j][0], acc[i][j][1], a[i], b[j]);
}

// ragged sub-tile: run-time strides
template<int MT, int NT>
__device__ __forceinline__ void ComputeStageR(double (&acc)[8][4][2], const FragAddrR &f) {
#pragma unroll
  for (int ks = 0; ks < RBK / 4; ++ks) KStepR<MT, NT>(acc, f.a + ks * f.a_k, f.a_i, f.b + (uint32_t(ks) ^ f.b_x) * f.b_k, f.b_j);
}

// full sub-tile, operand layouts known at compile time: every fragment load is a base register + immediate
template<bool AT, bool BT>
__device__ __forceinline__ void ComputeFullStageR(double (&acc)[8][4][2], const double *tileA, const double *tileB, int q, int g4, int t4) {
  constexpr uint32_t a_i = AT ? 8 : 8 * RLDA, a_k = AT ? 4 * RLDAT : 4;
  constexpr uint32_t b_j = BT ? 32 * RBK : 32, b_k = BT ? 4 : 4 * RLDB;
  const double *pa = tileA + (AT ? t4 * RLDAT + g4 : g4 * RLDA + t4);
  const double *pb = tileB + (BT ? (q * 8 + g4) * RBK + t4 : t4 * RLDB + q * 8 + g4);
  // transposed B: k4 group ks of row n sits at (ks ^ (n & 3)) -- one lane-constant base per step
  const int x = BT ? (g4 & 3) : 0;
#pragma unroll
  for (int ks = 0; ks < RBK / 4; ++ks) KStepR<8, 4>(acc, pa + ks * a_k, a_i, pb + (BT ? ((ks ^ x) * 4) : ks * int(b_k)), b_j);
}

__device__ __forceinline__ void NegateAccR(double (&acc)[8][4][2]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = -acc[i][j][0]; acc[i][j][1] = -acc[i][j][1]; }
}

template<int MT>
__device__ __forceinline__ void ComputeStageRN(double (&acc)[8][4][2], const FragAddrR &f, int nt) {
  switch (nt) {
    case 4: ComputeStageR<MT, 4>(acc, f); break;
    case 3: ComputeStageR<MT, 3>(acc, f); break;
    case 2: ComputeStageR<MT, 2>(acc, f); break;
    case 1: ComputeStageR<MT, 1>(acc, f); break;
    default: break;
  }
}

// Split-K fix-up (see gemm_ws.cu FixupTile): out of line, sums the partial tiles in slot order, writes C.
template<bool ACC>
__device__ __noinline__ void FixupTileR(const GemmParams &p, const GemmTile &tile, const GemmGroup &g, int q, int g4, int t4) {
  __threadfence();
  const uint32_t row0 = g.row_begin + uint32_t(tile.tm) * RBM, col0 = uint32_t(tile.tn) * RBN;
  const double *src0 = static_cast<const double *>(p.partials) + (unsigned long long) tile.part_base * (RBM * RBN) +
                       g4 * RBN + q * 8 + 2 * t4;
#pragma unroll 4
  for (int i = 0; i < 8; ++i) {
    double2 sum[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) sum[j] = make_double2(0.0, 0.0);
    const double *src = src0 + i * 8 * RBN;
    for (uint32_t sp = 0; sp < tile.nsplit; ++sp, src += RBM * RBN) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double2 v = __ldcg(reinterpret_cast<const double2 *>(src + j * 32));
        sum[j].x += v.x; sum[j].y += v.y;
      }
    }
    const uint32_t row = row0 + i * 8 + g4;
    if (row < g.row_end) {
      for (uint32_t d = 0; d < p.n_out; ++d) {
        double *Cg = static_cast<double *>(p.c_out[d]) + g.c_off + (unsigned long long) row * g.n;
        const double *Ci = static_cast<const double *>(p.c_in) + g.c_in_off + (unsigned long long) row * g.n;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t col = col0 + (q + 4 * j) * 8 + 2 * t4;
          if constexpr (ACC) {
            if (col < g.n) sum[j].x = AxpbyOut(p, sum[j].x, Ci + col, g.beta_on != 0);
            if (col + 1 < g.n) sum[j].y = AxpbyOut(p, sum[j].y, Ci + col + 1, g.beta_on != 0);
          }
          if (col + 1 < g.n && (reinterpret_cast<unsigned long long>(Cg + col) & 15ull) == 0) {
            StoreOut(reinterpret_cast<double2 *>(Cg + col), sum[j], p.mcast);
          } else {
            if (col < g.n) StoreOut(Cg + col, sum[j].x, p.mcast);
            if (col + 1 < g.n) StoreOut(Cg + col + 1, sum[j].y, p.mcast);
          }
        }
      }
    }
  }
  if (q == 0 && g4 == 0 && t4 == 0) p.counters[2 + tile.ctr] = 0;
}

template<bool ACC>
__global__ void __launch_bounds__(kWsThreads, 2)
GemmWsReal(const __grid_constant__ GemmParams p) {
  constexpr int STAGES = kRealStages;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *stages = reinterpret_cast<double *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + size_t(STAGES) * RSTAGE_ELEMS * sizeof(double));
  uint64_t *empty = full + STAGES;
  StageMeta *meta = reinterpret_cast<StageMeta *>(empty + STAGES);
  __shared__ uint32_t s_tile[2];      // tile id handed from producer warp 0 to the other producer warps
  __shared__ uint32_t s_last;         // split-K: this CTA holds the last unit of its tile

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { MbarInit(&full[s], kFullArrivals); MbarInit(&empty[s], kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= kConsumerWarps) {
    // ================================ producer warpgroup ================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    const uint32_t pw = warp - kConsumerWarps;     // each producer warp issues a quarter of a stage's copies
    const uint32_t a_kc = lane & 15, a_r = lane >> 4;
    uint32_t it = 0, tcount = 0;
    for (;; ++tcount) {
      if (pw == 0 && lane == 0) {
        uint32_t id;
        if (p.seg != nullptr) {      // stream-K: this CTA owns the units [seg[b], seg[b+1])
          const uint32_t u = p.seg[blockIdx.x] + tcount;
          id = u < p.seg[blockIdx.x + 1] ? u : p.ntiles;
        } else {
          id = atomicAdd(&p.counters[0], 1u);
        }
        s_tile[tcount & 1u] = id;
      }
      ProducerBarrier();
      const uint32_t tile_id = s_tile[tcount & 1u];
      if (tile_id >= p.ntiles) break;
      const GemmTile tile = p.tiles[tile_id];
      const GemmGroup g = p.groups[tile.group];
      const uint32_t row0 = g.row_begin + uint32_t(tile.tm) * RBM, col0 = uint32_t(tile.tn) * RBN;
      const uint32_t rows = min(uint32_t(RBM), g.row_end - row0), cols = min(uint32_t(RBN), g.n - col0);
      const uint32_t extents = (((rows + 7u) >> 3) << 8) | (((cols + 7u) >> 3) << 16);
      uint32_t sidx = 0;     // stage index of the current pair's first stage in the group's concatenated k loop
      for (uint32_t t = g.task_begin; t < g.task_end && sidx < tile.s_end; ++t) {
        const GemmTask task = p.tasks[t];
        const int next_sign = t + 1 < g.task_end ? p.tasks[t + 1].sign : 0;
        const uint32_t nst = (task.k + RBK - 1) / RBK;
        const uint32_t st_lo = max(sidx, tile.s_begin), st_hi = min(sidx + nst, tile.s_end);
        const uint32_t st_base = sidx;
        sidx += nst;
        if (st_lo >= st_hi) continue;
        const double *aBase = static_cast<const double *>((task.flags & kTaskASrc) ? p.a_src : p.a_ws) + task.a_off;
        const double *bBase = static_cast<const double *>((task.flags & kTaskBSrc) ? p.b_src : p.b_ws) + task.b_off;
        const bool ta = (task.flags & kTaskATrans) != 0, tb = (task.flags & kTaskBTrans) != 0;
        const uint32_t tflags = extents | (ta ? kFlagATrans : 0u) | (tb ? kFlagBTrans : 0u);
        // accumulator sign frame: flip after this pair's last stage if the unit goes on with a pair of the other sign, or
        // ends here in a negative frame
        const bool flip_after = st_hi == tile.s_end ? task.sign < 0 : (task.sign < 0) != (next_sign < 0);
        // B as a k x n view of the stored block (GemmTask::b_rs / b_cs / b_run): offset of this lane's four columns inside a row
        uint32_t bcol[4];
#pragma unroll
        for (uint32_t c = 0; c < 4; ++c) {
          const uint32_t colg = col0 + lane + 32u * c;
          bcol[c] = task.b_run >= g.n ? colg : (colg / task.b_run) * task.b_cs + colg % task.b_run;
        }
        for (uint32_t st = st_lo, k0 = (st_lo - st_base) * RBK; st < st_hi; ++st, ++it, k0 += RBK) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
          MbarWait(&empty[s], ph ^ 1u);
          const uint32_t sA = SmemAddr(stages + size_t(s) * RSTAGE_ELEMS);
          const uint32_t sB = sA + RA_ELEMS * 8u;
          if (!ta) {   // A row-major m x k: a lane copies element (a_r + 2r, a_kc) of the 64 x 16 tile
            const uint32_t kk = k0 + a_kc;
            const bool kok = kk < task.k;
            const double *src = aBase + (unsigned long long) (row0 + a_r) * task.k + kk;
#pragma unroll
            for (uint32_t rr = 0; rr < 8; ++rr) {
              const uint32_t r = 8u * pw + rr, row = a_r + 2u * r;
              const bool ok = kok && row < rows;
              CpAsync8Z(sA + (row * RLDA + a_kc) * 8u, ok ? src + (unsigned long long) (2u * r) * task.k : aBase, ok);
            }
#if QLB200_REAL_PF_STAGES > 0
            if (a_kc == 0 && k0 + uint32_t(QLB200_REAL_PF_STAGES) * RBK < task.k) {      // one lane per row, see the top of the file
#pragma unroll
              for (uint32_t rr = 0; rr < 8; ++rr) {
                const uint32_t r = 8u * pw + rr, row = a_r + 2u * r;
                if (row < rows) PrefetchL2(src + (unsigned long long) (2u * r) * task.k + uint32_t(QLB200_REAL_PF_STAGES) * RBK);
              }
            }
#endif
          } else {     // A stored k x m: 16 k-rows of 64 contiguous elements
            const double *src = aBase + (unsigned long long) k0 * g.m + row0 + lane;
#pragma unroll
            for (uint32_t rr = 0; rr < 4; ++rr) {
              const uint32_t kr = 4u * pw + rr;
              const bool kok = k0 + kr < task.k;
#pragma unroll
              for (uint32_t c = 0; c < 2; ++c) {
                const uint32_t ml = lane + 32u * c;
                const bool ok = kok && ml < rows;
                CpAsync8Z(sA + (kr * RLDAT + ml) * 8u, ok ? src + (unsigned long long) kr * g.m + 32u * c : aBase, ok);
              }
            }
          }
          if (!tb) {   // B row-major k x n: 16 k-rows x 128 columns, 256 contiguous bytes per copy
#pragma unroll
            for (uint32_t rr = 0; rr < 4; ++rr) {
              const uint32_t kr = 4u * pw + rr;
              const bool rok = k0 + kr < task.k;
              const double *src = bBase + (unsigned long long) (k0 + kr) * task.b_rs;
#pragma unroll
              for (uint32_t c = 0; c < 4; ++c) {
                const uint32_t col = lane + 32u * c;
                const bool ok = rok && col < cols;
                CpAsync8Z(sB + (kr * RLDB + col) * 8u, ok ? src + bcol[c] : bBase, ok);
              }
            }
          } else {     // B stored n x k: a lane copies element (a_r + 2r, a_kc) of the 128 x 16 tile
            const uint32_t kk = k0 + a_kc;
            const bool kok = kk < task.k;
            const double *src = bBase + (unsigned long long) (col0 + a_r) * task.k + kk;
#pragma unroll
            for (uint32_t rr = 0; rr < 16; ++rr) {
              const uint32_t r = 16u * pw + rr, nl = a_r + 2u * r;
              const bool ok = kok && nl < cols;
              CpAsync8Z(sB + (nl * RBK + (a_kc ^ ((nl & 3u) << 2))) * 8u, ok ? src + (unsigned long long) (2u * r) * task.k : bBase, ok);
            }
          }
          CpAsyncMbarArrive(&full[s]);
          if (pw == 0 && lane == 0) {
            uint32_t fl = tflags;
            if (st == tile.s_begin) fl |= kFlagFirst;
            if (st + 1 == tile.s_end) fl |= kFlagLast;
            if (st + 1 == st_hi && flip_after) fl |= kFlagFlip;
            meta[s].tile = tile_id; meta[s].flags = fl;
            MbarArrive(&full[s]);
          }
        }
      }
    }
    {   // sentinel stage: tells every consumer warp to leave
      const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
      MbarWait(&empty[s], ph ^ 1u);
      CpAsyncMbarArrive(&full[s]);
      if (pw == 0 && lane == 0) { meta[s].tile = kSentinel; meta[s].flags = 0; MbarArrive(&full[s]); }
    }
    if (pw == 0 && lane == 0) {
      __threadfence();
      if (atomicAdd(&p.counters[1], 1u) == gridDim.x - 1) { p.counters[0] = 0; p.counters[1] = 0; __threadfence(); }
    }
    return;
  }

  // ==================================== consumer warps ====================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
  const int q = warp;
  const int g4 = lane >> 2, t4 = lane & 3;
  double acc[8][4][2];
  uint32_t s = 0, ph = 0;     // ring position and phase parity
  for (;; ph ^= (++s == uint32_t(STAGES)) ? 1u : 0u, s = (s == uint32_t(STAGES)) ? 0u : s) {
    MbarWait(&full[s], ph);
    const StageMeta sm = meta[s];
    if (sm.tile == kSentinel) break;
    if (sm.flags & kFlagFirst) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    }
    const int mt = int((sm.flags >> 8) & 0xfu);
    const int n8 = int((sm.flags >> 16) & 0x1fu);
    const int nt = n8 > q ? (n8 - q + 3) >> 2 : 0;
    const double *tileA = stages + size_t(s) * RSTAGE_ELEMS;
    const double *tileB = tileA + RA_ELEMS;
    if (mt == 8 && nt == 4) {
      switch (sm.flags & (kFlagATrans | kFlagBTrans)) {
        case 0: ComputeFullStageR<false, false>(acc, tileA, tileB, q, g4, t4); break;
        case kFlagATrans: ComputeFullStageR<true, false>(acc, tileA, tileB, q, g4, t4); break;
        case kFlagBTrans: ComputeFullStageR<false, true>(acc, tileA, tileB, q, g4, t4); break;
        default: ComputeFullStageR<true, true>(acc, tileA, tileB, q, g4, t4); break;
      }
    } else {
      FragAddrR f;
      if (sm.flags & kFlagATrans) { f.a = tileA + t4 * RLDAT + g4; f.a_i = 8; f.a_k = 4 * RLDAT; }
      else { f.a = tileA + g4 * RLDA + t4; f.a_i = 8 * RLDA; f.a_k = 4; }
      if (sm.flags & kFlagBTrans) { f.b = tileB + (q * 8 + g4) * RBK + t4; f.b_j = 32 * RBK; f.b_k = 4; f.b_x = g4 & 3; }
      else { f.b = tileB + t4 * RLDB + q * 8 + g4; f.b_j = 32; f.b_k = 4 * RLDB; f.b_x = 0; }
      switch ((mt + 1) >> 1) {      // m8 groups are specialised in pairs
        case 4: ComputeStageRN<8>(acc, f, nt); break;
        case 3: ComputeStageRN<6>(acc, f, nt); break;
        case 2: ComputeStageRN<4>(acc, f, nt); break;
        default: ComputeStageRN<2>(acc, f, nt); break;
      }
    }
    __syncwarp();
    if (lane == 0) MbarArrive(&empty[s]);
    if (sm.flags & kFlagFlip) NegateAccR(acc);      // sign frame change (producer-marked)
    if (sm.flags & kFlagLast) {
      const GemmTile tile = p.tiles[sm.tile];
      const GemmGroup g = p.groups[tile.group];
      const uint32_t row0 = g.row_begin + uint32_t(tile.tm) * RBM, col0 = uint32_t(tile.tn) * RBN;
      bool write_c = true;
      if (tile.nsplit > 1) {
        // deterministic split-K: park this unit's partial tile, the last unit to arrive adds all of them
        double *slots = static_cast<double *>(p.partials) + (unsigned long long) tile.part_base * (RBM * RBN);
        double *mine = slots + (unsigned long long) tile.split * (RBM * RBN) + g4 * RBN + q * 8 + 2 * t4;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<double2 *>(mine + i * 8 * RBN + j * 32) = make_double2(acc[i][j][0], acc[i][j][1]);
        __threadfence();
        ConsumerBarrier();
        if (warp == 0 && lane == 0) s_last = atomicAdd(&p.counters[2 + tile.ctr], 1u) == uint32_t(tile.nsplit) - 1u ? 1u : 0u;
        ConsumerBarrier();
        write_c = false;
        if (s_last != 0) FixupTileR<ACC>(p, tile, g, q, g4, t4);
      }
#ifdef QLB200_EXP_NOEPI
      if (acc[0][0][0] != 12345.678) write_c = false;
#endif
      if (write_c) {
        for (uint32_t d = 0; d < p.n_out; ++d) {     // n_out > 1: fused exchange, the same tile goes to every NVLink peer
          double *Cg = static_cast<double *>(p.c_out[d]) + g.c_off;
          const double *Ci = static_cast<const double *>(p.c_in) + g.c_in_off;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t row = row0 + i * 8 + g4;
            if (row >= g.row_end) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t col = col0 + (q + 4 * j) * 8 + 2 * t4;
              const unsigned long long at = (unsigned long long) row * g.n + col;
              double *dst = Cg + at;
              if constexpr (ACC) {
                if (col < g.n) StoreOut(dst, AxpbyOut(p, acc[i][j][0], Ci + at, g.beta_on != 0), p.mcast);
                if (col + 1 < g.n) StoreOut(dst + 1, AxpbyOut(p, acc[i][j][1], Ci + at + 1, g.beta_on != 0), p.mcast);
              } else if (col + 1 < g.n && (reinterpret_cast<unsigned long long>(dst) & 15ull) == 0) {
                // the lane's two adjacent columns as one 16-byte store where the row happens to be aligned (half as many store
                // instructions per tile; rows of odd length alternate)
                StoreOut(reinterpret_cast<double2 *>(dst), make_double2(acc[i][j][0], acc[i][j][1]), p.mcast);
              } else {
                if (col < g.n) StoreOut(dst, acc[i][j][0], p.mcast);
                if (col + 1 < g.n) StoreOut(dst + 1, acc[i][j][1], p.mcast);
              }
            }
          }
        }
      }
    }
  }
}

}  // namespace

cudaError_t ConfigureWsRealKernel() {
  cudaError_t e = cudaFuncSetAttribute(GemmWsReal<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kRealWsSmem));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(GemmWsReal<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kRealWsSmem));
}

cudaError_t LaunchGemmWsReal(const GemmParams &p, int num_sms, cudaStream_t stream) {
  if (p.ntiles == 0) return cudaSuccess;
  const uint32_t cap = 2u * uint32_t(num_sms);     // two resident CTAs per SM
  const uint32_t grid = p.seg != nullptr ? p.nseg : (p.ntiles < cap ? p.ntiles : cap);
  if (p.accum) GemmWsReal<true><<<grid, kWsThreads, kRealWsSmem, stream>>>(p);
  else GemmWsReal<false><<<grid, kWsThreads, kRealWsSmem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace qlb200
