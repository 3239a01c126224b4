// gemm_ws.cu -- warp-specialised grouped complex-FP64 GEMM (the headline kernel).
//
// Same contract as gemm.cu (C = sum over the group's pairs of sign * A[m x k] * B[k x n], written
// once, deterministic), different machine mapping:
//
//   * CTA = 4 consumer warps + 4 producer warps (registers re-split with setmaxnreg), CTA tile 32 x 128,
//     TWO CTAs resident per SM.  The two
//     CTAs of an SM drift apart, so one CTA's epilogue / prologue is covered by the other CTA's main
//     loop and the FP64 tensor pipe (DMMA.8x8x4) of every SM sub-partition always has a warp to feed it.
//   * The producer warps move operand tiles global -> shared with 16-byte asynchronous copies
//     (cp.async, SASS LDGSTS; complex elements are 16 bytes, so any block offset / leading dimension
//     is legal) and signals completion through an mbarrier (cp.async.mbarrier.arrive).  Ragged edges
//     and the K tail are zero-filled by the copy itself (src-size 0).
//   * Stages are handed over with full/empty mbarriers only -- no CTA-wide barrier in the main loop.
//   * Every stage carries {tile id, first/last/sign flags, valid extents}; the producer runs ahead
//     across tile boundaries (dynamic atomic tile counter), so the next tile's prologue overlaps this
//     tile's epilogue stores.
//   * Operand blocks are read where they lie: row-major (A: m x k, B: k x n) or 2-D transposed
//     (A stored k x m, B stored n x k; GemmTask::flags) -- the producer picks the copy pattern that
//     keeps global reads contiguous, the consumers only change their shared-memory strides.  Most
//     blocks of a DMRG contraction therefore never pass through the permute kernel.
//   * Ragged tiles cost what they use: a warp owns the n8 column groups {q, q+4, q+8, q+12} of the
//     tile (interleaved, so valid columns spread evenly over the four sub-partitions) and skips the
//     MMAs of m8 row groups / n8 column groups that lie outside the output block.
//   * Two arithmetic schemes (template parameter CFG).  Cfg4M is the textbook product: four real DMMAs
//     per complex 8x8x4 step, CTA tile 32 x 128.  Cfg3M (default) is the Gauss / Karatsuba product
//         P1 += Ar Br,   P2 += Ai Bi,   P3 += (Ar + Ai) (Br + Bi);   Cr = P1 - P2,  Ci = P3 - P1 - P2
//     -- three real DMMAs per step (the form with the fewest operand pre-additions: one per A and one
//     per B fragment; they run on the same FP64 datapath as DMMA and cost ~5 pipe cycles each).  The kernel sits at ~90 % DMMA-pipe utilisation (ncu), so the only
//     way to go faster is to issue fewer DMMAs: 3M cuts them by 25 %.  The three running sums need 3/2
//     the accumulator registers, hence the narrower 32 x 96 CTA tile (a warp owns 3 n8 groups); the
//     operand sums cost 7 DADDs per 36 DMMAs.  Deterministic like 4M; normwise error bound of the same
//     order (the parity bar is a relative Frobenius error <= 1e-12, tests/test_parity_gpu.py).
#include "common.cuh"
#include "ws_common.cuh"
#ifdef QLB200_EXP_NOCOPY
#define CpAsync16Z(a, b, c) ((void) 0)
#endif
#ifndef QLB200_WS_FULLSTAGE
#define QLB200_WS_FULLSTAGE 1
#endif


namespace qlb200 {

namespace {

constexpr int WBM = kWsBM;
struct Cfg4M { static constexpr int BN = kWsBN, BK = kWsBK, NT = kWsBN / 32, NACC = 2, STAGES = 5; static constexpr bool k3M = false; };
struct Cfg3M { static constexpr int BN = kWs3mBN, BK = kWs3mBK, NT = kWs3mBN / 32, NACC = 3, STAGES = kWs3mStages; static constexpr bool k3M = true; };
// Shared-memory tile layouts (units: complex elements = one 16-byte bank group); every fragment load
// of a quarter-warp (lanes g4 in {2p, 2p+1}, t4 in 0..3) hits 8 distinct bank groups:
//   A row-major   [32 m][BK + 4]  ((BK+4)*g4 + t4) mod 8 = 4*(g4&1) + t4 distinct   (BK = 8 or 16)
//   A transposed  [BK k][34]      (34*t4 + g4)  mod 8 = 2*t4 + g4 distinct
//   B row-major   [BK k][BN + 2]  ((BN+2)*t4 + g4) mod 8 = 2*t4 + g4 distinct   (BN = 128 or 96)
//   B transposed  [BN n][BK] with adjacent k4 groups of odd rows swapped (k ^ 4*(n&1)): 4*(g4&1) + t4 distinct
constexpr int WLDAT = WBM + 2;

template<class CFG>
struct Lay {
  static constexpr int WBN = CFG::BN, WBK = CFG::BK, WLDA = WBK + 4, WLDB = WBN + 2;
  static constexpr int A_ELEMS = WBM * WLDA, B_ELEMS = WBK * WLDB, STAGE_ELEMS = A_ELEMS + B_ELEMS;
  static_assert(WBK * WLDAT <= A_ELEMS, "transposed A tile must fit the stage");
  static_assert(WBN * WBK <= B_ELEMS, "transposed B tile must fit the stage");
  static_assert(WLDB % 8 == 2 && WLDA % 8 == 4, "bank-group spread of the fragment loads");
};

template<class CFG, int STAGES>
struct WsSmem {
  static constexpr size_t kBytes = size_t(STAGES) * Lay<CFG>::STAGE_ELEMS * sizeof(double2) + 2 * STAGES * sizeof(uint64_t) +
                                   STAGES * sizeof(StageMeta);
};

// Fragment addressing of one stage (element units, relative to the stage's A / B regions).
struct FragAddr {
  const double2 *a, *b;     // lane's base inside the A / B tile
  uint32_t a_i, a_ks;       // A: + i * a_i (m8 group) + ks * a_ks (k4 step)
  uint32_t b_j, b_ks;       // B: + j * b_j (owned n8 group) + ks * b_ks  (row-major B; b_sw = 0)
  uint32_t b_sw, b_x;       //    + ((ks ^ b_x) * b_sw)                   (transposed B: swizzled k4 groups; b_ks = 0)
};

// One k4 step of a warp's sub-tile: MT valid m8 row groups x NT valid n8 column groups, fragments at pa + i * a_i and
// pb + j * b_j.  Specialised at compile time so that skipped MMAs are not even issued; with compile-time strides (the
// full-tile path below) every fragment load is a base register + immediate.
// acc[0] / acc[1] = real / imaginary sums (4M);  acc[0..2] = P1, P2, P3 (3M).
// The pair's sign is NOT applied here: the consumer keeps the accumulators in the sign frame of the current pair
// (NegateAcc on a change of sign, bit-identical to negating the A fragments: fma(-a, b, c) = -fma(a, b, -c)).
template<class CFG, int MT, int NT>
__device__ __forceinline__ void KStep(double (&acc)[CFG::NACC][4][CFG::NT][2], const double2 *pa, uint32_t a_i, const double2 *pb, uint32_t b_j) {
  if constexpr (CFG::k3M) {
    double as[MT], ai[MT], ar[MT];
    double br[NT], bs[NT], bi[NT];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      const double2 a = pa[i * a_i];
      ar[i] = a.x; ai[i] = a.y; as[i] = a.x + a.y;
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const double2 b = pb[j * b_j];
      br[j] = b.x; bi[j] = b.y; bs[j] = b.x + b.y;
    }
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) DmmaNv(acc[0][i][j][0], acc[0][i][j][1], ar[i], br[j]);
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) DmmaNv(acc[1][i][j][0], acc[1][i][j][1], ai[i], bi[j]);
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) DmmaNv(acc[2][i][j][0], acc[2][i][j][1], as[i], bs[j]);
  } else {
    double ax[MT], ay[MT], nay[MT];
    double2 b[NT];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      const double2 a = pa[i * a_i];
      ax[i] = a.x; ay[i] = a.y; nay[i] = FlipSign(a.y, 0x80000000u);
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) b[j] = pb[j * b_j];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        DmmaNv(acc[0][i][j][0], acc[0][i][j][1], ax[i], b[j].x);
        DmmaNv(acc[1][i][j][0], acc[1][i][j][1], ax[i], b[j].y);
      }
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        DmmaNv(acc[0][i][j][0], acc[0][i][j][1], nay[i], b[j].y);
        DmmaNv(acc[1][i][j][0], acc[1][i][j][1], ay[i], b[j].x);
      }
  }
}

// Two k4 steps (k4 groups KS0, KS0+1 of the stage) of a ragged sub-tile, run-time strides.
template<class CFG, int MT, int NT, int KS0>
__device__ __forceinline__ void ComputeStage(double (&acc)[CFG::NACC][4][CFG::NT][2], const FragAddr &f) {
#pragma unroll
  for (int ks = KS0; ks < KS0 + 2; ++ks)
    KStep<CFG, MT, NT>(acc, f.a + ks * f.a_ks, f.a_i, f.b + ks * f.b_ks + ((uint32_t(ks) ^ f.b_x) * f.b_sw), f.b_j);
}

// A whole stage (CFG::BK / 4 k4 steps) of a full sub-tile with the operand layouts known at compile time: the common case
// (interior tiles of large blocks), one straight-line run of MMAs per stage, no address arithmetic between them.
template<class CFG, bool AT, bool BT>
__device__ __forceinline__ void ComputeFullStage(double (&acc)[CFG::NACC][4][CFG::NT][2], const double2 *tileA, const double2 *tileB,
                                                 int q, int g4, int t4) {
  constexpr int WBK = CFG::BK, WLDA = Lay<CFG>::WLDA, WLDB = Lay<CFG>::WLDB;
  constexpr uint32_t a_i = AT ? 8 : 8 * WLDA, a_ks = AT ? 4 * WLDAT : 4;
  constexpr uint32_t b_j = BT ? 32 * WBK : 32, b_ks = BT ? 4 : 4 * WLDB;
  const double2 *pa = tileA + (AT ? t4 * WLDAT + g4 : g4 * WLDA + t4);
  const double2 *pb = tileB + (BT ? (q * 8 + g4) * WBK + t4 : t4 * WLDB + q * 8 + g4);
  // transposed B: the k4 groups of odd rows are swapped in pairs (k ^ 4 * (n & 1)); even steps then sit 4 * (g4 & 1) elements
  // further, odd steps as many elements earlier -- two lane-constant bases, immediates for everything else
  const int sw = BT ? 4 * (g4 & 1) : 0;
  const double2 *pb_even = pb + sw, *pb_odd = pb - sw;
#pragma unroll
  for (int ks = 0; ks < WBK / 4; ++ks)
    KStep<CFG, 4, CFG::NT>(acc, pa + ks * a_ks, a_i, ((ks & 1) ? pb_odd : pb_even) + ks * b_ks, b_j);
}

template<class CFG>
__device__ __forceinline__ void NegateAcc(double (&acc)[CFG::NACC][4][CFG::NT][2]) {
#pragma unroll
  for (int a = 0; a < CFG::NACC; ++a)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < CFG::NT; ++j) {
        acc[a][i][j][0] = -acc[a][i][j][0];
        acc[a][i][j][1] = -acc[a][i][j][1];
      }
}

template<class CFG, int MT, int KS0>
__device__ __forceinline__ void ComputeStageN(double (&acc)[CFG::NACC][4][CFG::NT][2], const FragAddr &f, int nt) {
  if constexpr (CFG::NT >= 4) { if (nt == 4) { ComputeStage<CFG, MT, 4, KS0>(acc, f); return; } }
  switch (nt) {
    case 3: ComputeStage<CFG, MT, 3, KS0>(acc, f); break;
    case 2: ComputeStage<CFG, MT, 2, KS0>(acc, f); break;
    case 1: ComputeStage<CFG, MT, 1, KS0>(acc, f); break;
    default: break;
  }
}

template<class CFG, int KS0>
__device__ __forceinline__ void ComputeHalf(double (&acc)[CFG::NACC][4][CFG::NT][2], const FragAddr &f, int mt, int nt) {
  if (mt == 4 && nt == CFG::NT) {
    ComputeStage<CFG, 4, CFG::NT, KS0>(acc, f);
  } else {
    switch (mt) {
      case 4: ComputeStageN<CFG, 4, KS0>(acc, f, nt); break;
      case 3: ComputeStageN<CFG, 3, KS0>(acc, f, nt); break;
      case 2: ComputeStageN<CFG, 2, KS0>(acc, f, nt); break;
      default: ComputeStageN<CFG, 1, KS0>(acc, f, nt); break;
    }
  }
}

// complex value of accumulator element (i, j, e)
template<class CFG>
__device__ __forceinline__ double2 AccValue(const double (&acc)[CFG::NACC][4][CFG::NT][2], int i, int j, int e) {
  if constexpr (CFG::k3M) return make_double2(acc[0][i][j][e] - acc[1][i][j][e], (acc[2][i][j][e] - acc[0][i][j][e]) - acc[1][i][j][e]);
  else return make_double2(acc[0][i][j][e], acc[1][i][j][e]);
}

// Split-K fix-up, run by the unit that arrived last: add the tile's partial tiles in slot order (a fixed
// order, whichever unit happens to be last) one m8 row group at a time and write C.  Kept out of line:
// it needs only a handful of registers and must not disturb the register allocation of the main loop.
template<class CFG, bool MCAST, bool ACC>
__device__ __noinline__ void FixupTile(const GemmParams &p, const GemmTile &tile, const GemmGroup &g, int q, int g4, int t4) {
  constexpr int WBN = CFG::BN, NT = CFG::NT;
  __threadfence();
  const uint32_t row0 = g.row_begin + uint32_t(tile.tm) * WBM, col0 = uint32_t(tile.tn) * WBN;
  const double2 *src0 = static_cast<const double2 *>(p.partials) + (unsigned long long) tile.part_base * (WBM * WBN) +
                        g4 * WBN + q * 8 + 2 * t4;
#pragma unroll 2
  for (int i = 0; i < 4; ++i) {
    double2 sum[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) sum[j][0] = sum[j][1] = make_double2(0.0, 0.0);
    const double2 *src = src0 + i * 8 * WBN;
    for (uint32_t sp = 0; sp < tile.nsplit; ++sp, src += WBM * WBN) {
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const double2 v0 = __ldcg(src + j * 32), v1 = __ldcg(src + j * 32 + 1);
        sum[j][0].x += v0.x; sum[j][0].y += v0.y; sum[j][1].x += v1.x; sum[j][1].y += v1.y;
      }
    }
    const uint32_t row = row0 + i * 8 + g4;
    if (row < g.row_end) {
      for (uint32_t d = 0; d < p.n_out; ++d) {
        double2 *Cg = static_cast<double2 *>(p.c_out[d]) + g.c_off + (unsigned long long) row * g.n;
        const double2 *Ci = static_cast<const double2 *>(p.c_in) + g.c_in_off + (unsigned long long) row * g.n;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const uint32_t col = col0 + (q + 4 * j) * 8 + 2 * t4;
          if constexpr (ACC) {
            if (col < g.n) sum[j][0] = AxpbyOut(p, sum[j][0], Ci + col, g.beta_on != 0);
            if (col + 1 < g.n) sum[j][1] = AxpbyOut(p, sum[j][1], Ci + col + 1, g.beta_on != 0);
          }
          if (col < g.n) StoreOut(Cg + col, sum[j][0], MCAST);
          if (col + 1 < g.n) StoreOut(Cg + col + 1, sum[j][1], MCAST);
        }
      }
    }
  }
  if (q == 0 && g4 == 0 && t4 == 0) p.counters[2 + tile.ctr] = 0;
}

// Accumulate epilogue of one tile row, out of line (like FixupTile: it must not disturb the register allocation of the main
// loop, which has no register to spare next to the 3M accumulators): out = beta * C_in + alpha * v for the lane's up to
// eight elements of the row; column of element (j, e) = col + 32 * j + e.
__device__ __noinline__ void StoreAccRow(const GemmParams &p, double2 *crow, const double2 *cin_row, uint32_t col, uint32_t n, uint32_t beta_on,
                                         int nt, double2 v00, double2 v01, double2 v10, double2 v11, double2 v20, double2 v21, double2 v30,
                                         double2 v31) {
  const double2 v[4][2] = {{v00, v01}, {v10, v11}, {v20, v21}, {v30, v31}};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j >= nt) break;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const uint32_t c = col + 32u * j + e;
      if (c < n) crow[c] = AxpbyOut(p, v[j][e], cin_row + c, beta_on != 0);
    }
  }
}

// MCAST: the output address is an NVSwitch multicast mapping (multimem.st); a compile-time switch, because a
// run-time branch around every store of the unrolled epilogue costs registers the main loop cannot spare.
// ACC: accumulate form, the epilogue computes beta * C_in + alpha * (sum of pairs) (qlb200_execute_accum); also compile-time.
template<class CFG, int STAGES, bool MCAST, bool ACC>
__global__ void __launch_bounds__(kWsThreads, 2)
GemmWsCplx(const __grid_constant__ GemmParams p) {
  constexpr int WBN = CFG::BN, WBK = CFG::BK, WLDA = Lay<CFG>::WLDA, WLDB = Lay<CFG>::WLDB, A_ELEMS = Lay<CFG>::A_ELEMS,
                STAGE_ELEMS = Lay<CFG>::STAGE_ELEMS, NTMAX = CFG::NT;
  constexpr uint32_t RP = 32 / WBK;      // rows of a [rows][WBK] tile one warp pass covers (a lane copies one element)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2 *stages = reinterpret_cast<double2 *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + size_t(STAGES) * STAGE_ELEMS * sizeof(double2));
  uint64_t *empty = full + STAGES;
  StageMeta *meta = reinterpret_cast<StageMeta *>(empty + STAGES);
  __shared__ uint32_t s_tile[2];      // tile id handed from producer warp 0 to the other producer warps
  __shared__ uint32_t s_last;         // split-K: this CTA holds the last unit of its tile

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    // full: one asynchronous copy-completion arrival per producer lane + one plain arrival (releases the stage meta)
    for (int s = 0; s < STAGES; ++s) { MbarInit(&full[s], kFullArrivals); MbarInit(&empty[s], kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= kConsumerWarps) {
    // ================================ producer warpgroup ================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    // the four producer warps walk the same tile / stage sequence and each issue a quarter of a stage's copies
    const uint32_t pw = warp - kConsumerWarps;
    const uint32_t a_kc = lane % WBK, a_r = lane / WBK;
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
      const uint32_t row0 = g.row_begin + uint32_t(tile.tm) * WBM, col0 = uint32_t(tile.tn) * WBN;
      const uint32_t rows = min(uint32_t(WBM), g.row_end - row0), cols = min(uint32_t(WBN), g.n - col0);
      const uint32_t extents = (((rows + 7u) >> 3) << 8) | (((cols + 7u) >> 3) << 16);
      uint32_t sidx = 0;     // stage index of the current pair's first stage in the group's concatenated k loop
      for (uint32_t t = g.task_begin; t < g.task_end && sidx < tile.s_end; ++t) {
        const GemmTask task = p.tasks[t];
        const int next_sign = t + 1 < g.task_end ? p.tasks[t + 1].sign : 0;
        const uint32_t nst = (task.k + WBK - 1) / WBK;
        const uint32_t st_lo = max(sidx, tile.s_begin), st_hi = min(sidx + nst, tile.s_end);
        const uint32_t st_base = sidx;
        sidx += nst;
        if (st_lo >= st_hi) continue;
        const double2 *aBase = static_cast<const double2 *>((task.flags & kTaskASrc) ? p.a_src : p.a_ws) + task.a_off;
        const double2 *bBase = static_cast<const double2 *>((task.flags & kTaskBSrc) ? p.b_src : p.b_ws) + task.b_off;
        const bool ta = (task.flags & kTaskATrans) != 0, tb = (task.flags & kTaskBTrans) != 0;
        const uint32_t tflags = extents | (ta ? kFlagATrans : 0u) | (tb ? kFlagBTrans : 0u);
        // accumulator sign frame (see the consumer): flip after this pair's last stage if the unit goes on with a pair of the
        // other sign, or ends here in a negative frame
        const bool flip_after = st_hi == tile.s_end ? task.sign < 0 : (task.sign < 0) != (next_sign < 0);
        // B as a k x n view of the stored block (GemmTask::b_rs / b_cs / b_run): offset of this lane's columns inside a row
        uint32_t bcol[WBN / 32];
#pragma unroll
        for (uint32_t c = 0; c < uint32_t(WBN / 32); ++c) {
          const uint32_t colg = col0 + lane + 32u * c;
          bcol[c] = task.b_run >= g.n ? colg : (colg / task.b_run) * task.b_cs + colg % task.b_run;
        }
        for (uint32_t st = st_lo, k0 = (st_lo - st_base) * WBK; st < st_hi; ++st, ++it, k0 += WBK) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
          MbarWait(&empty[s], ph ^ 1u);
          const uint32_t sA = SmemAddr(stages + size_t(s) * STAGE_ELEMS);
          const uint32_t sB = sA + A_ELEMS * 16u;
          if (!ta) {   // A row-major m x k: a lane copies element (a_r + RP*r, a_kc) of the 32 x WBK tile
            const uint32_t kk = k0 + a_kc;
            const bool kok = kk < task.k;
            const double2 *src = aBase + (unsigned long long) (row0 + a_r) * task.k + kk;
#pragma unroll
            for (uint32_t rr = 0; rr < WBM / RP / 4; ++rr) {
              const uint32_t r = (WBM / RP / 4) * pw + rr, row = a_r + RP * r;
              const bool ok = kok && row < rows;
              CpAsync16Z(sA + (row * WLDA + a_kc) * 16u, ok ? src + (unsigned long long) (RP * r) * task.k : aBase, ok);
            }
          } else {     // A stored k x m: WBK k-rows of 32 contiguous elements
            const double2 *src = aBase + (unsigned long long) k0 * g.m + row0 + lane;
            const bool mok = uint32_t(lane) < rows;
#pragma unroll
            for (uint32_t rr = 0; rr < WBK / 4; ++rr) {
              const uint32_t kr = (WBK / 4) * pw + rr;
              const bool ok = mok && k0 + kr < task.k;
              CpAsync16Z(sA + (kr * WLDAT + lane) * 16u, ok ? src + (unsigned long long) kr * g.m : aBase, ok);
            }
          }
          if (!tb) {   // B row-major k x n: WBK k-rows x WBN columns, 512 contiguous bytes per copy
#pragma unroll
            for (uint32_t rr = 0; rr < WBK / 4; ++rr) {
              const uint32_t kr = (WBK / 4) * pw + rr;
              const bool rok = k0 + kr < task.k;
              const double2 *src = bBase + (unsigned long long) (k0 + kr) * task.b_rs;
#pragma unroll
              for (uint32_t c = 0; c < uint32_t(WBN / 32); ++c) {
                const uint32_t col = lane + 32u * c;
                const bool ok = rok && col < cols;
                CpAsync16Z(sB + (kr * WLDB + col) * 16u, ok ? src + bcol[c] : bBase, ok);
              }
            }
          } else {     // B stored n x k: a lane copies element (a_r + RP*r, a_kc) of the WBN x WBK tile
            const uint32_t kk = k0 + a_kc;
            const bool kok = kk < task.k;
            const double2 *src = bBase + (unsigned long long) (col0 + a_r) * task.k + kk;
#pragma unroll
            for (uint32_t rr = 0; rr < WBN / RP / 4; ++rr) {
              const uint32_t r = (WBN / RP / 4) * pw + rr, nl = a_r + RP * r;
              const bool ok = kok && nl < cols;
              CpAsync16Z(sB + (nl * WBK + (a_kc ^ ((nl & 1u) << 2))) * 16u, ok ? src + (unsigned long long) (RP * r) * task.k : bBase, ok);
            }
          }
          CpAsyncMbarArrive(&full[s]);
          if (pw == 0 && lane == 0) {
            uint32_t fl = tflags | (min(uint32_t(WBK / 4), (task.k - k0 + 3u) >> 2) << 24);
            if (st == tile.s_begin) fl |= kFlagFirst;
            if (st + 1 == tile.s_end) fl |= kFlagLast;
            if (st + 1 == st_hi && flip_after) fl |= kFlagFlip;
            meta[s].tile = tile_id; meta[s].flags = fl;
            MbarArrive(&full[s]);
          }
        }
      }
    }
    // sentinel stage: tells every consumer warp to leave
    {
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
  // warp q owns all 32 rows and the n8 column groups {q, q+4, q+8, q+12} of the CTA tile
  const int q = warp;
  const int g4 = lane >> 2, t4 = lane & 3;
  double acc[CFG::NACC][4][NTMAX][2];
  uint32_t s = 0, ph = 0;     // ring position and phase parity
  for (;; ph ^= (++s == uint32_t(STAGES)) ? 1u : 0u, s = (s == uint32_t(STAGES)) ? 0u : s) {
    MbarWait(&full[s], ph);
    const StageMeta sm = meta[s];
    if (sm.tile == kSentinel) break;
    if (sm.flags & kFlagFirst) {
#pragma unroll
      for (int a = 0; a < CFG::NACC; ++a)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < NTMAX; ++j) acc[a][i][j][0] = acc[a][i][j][1] = 0.0;
    }
    const int mt = int((sm.flags >> 8) & 0xfu);
    const int n8 = int((sm.flags >> 16) & 0x1fu);
    const int nt = n8 > q ? (n8 - q + 3) >> 2 : 0;
    const double2 *tileA = stages + size_t(s) * STAGE_ELEMS;
    const double2 *tileB = tileA + A_ELEMS;
    const uint32_t nk4 = (sm.flags >> 24) & 7u;      // valid k4 steps of the stage (fewer than WBK / 4 only in a K tail)
    if (QLB200_WS_FULLSTAGE && mt == 4 && nt == NTMAX && nk4 == uint32_t(WBK / 4)) {
      switch (sm.flags & (kFlagATrans | kFlagBTrans)) {
        case 0: ComputeFullStage<CFG, false, false>(acc, tileA, tileB, q, g4, t4); break;
        case kFlagATrans: ComputeFullStage<CFG, true, false>(acc, tileA, tileB, q, g4, t4); break;
        case kFlagBTrans: ComputeFullStage<CFG, false, true>(acc, tileA, tileB, q, g4, t4); break;
        default: ComputeFullStage<CFG, true, true>(acc, tileA, tileB, q, g4, t4); break;
      }
    } else {
      FragAddr f;
      if (sm.flags & kFlagATrans) { f.a = tileA + t4 * WLDAT + g4; f.a_i = 8; f.a_ks = 4 * WLDAT; }
      else { f.a = tileA + g4 * WLDA + t4; f.a_i = 8 * WLDA; f.a_ks = 4; }
      if (sm.flags & kFlagBTrans) { f.b = tileB + (q * 8 + g4) * WBK + t4; f.b_j = 32 * WBK; f.b_ks = 0; f.b_sw = 4; f.b_x = g4 & 1; }
      else { f.b = tileB + t4 * WLDB + q * 8 + g4; f.b_j = 32; f.b_ks = 4 * WLDB; f.b_sw = 0; f.b_x = 0; }
      ComputeHalf<CFG, 0>(acc, f, mt, nt);
      if constexpr (WBK == 16) {
        if (nk4 > 2u) ComputeHalf<CFG, 2>(acc, f, mt, nt);   // K tail: skip an all-zero half stage
      }
    }
    __syncwarp();
    if (lane == 0) MbarArrive(&empty[s]);
    // Sign frame: the accumulators hold (sign of the current pair) x (sum so far).  The producer marks the stage after which
    // the frame changes (the next pair has the other sign, or the unit ends in a negative frame): one flip of the
    // accumulators there instead of a flip of every A fragment of every k step.
    if (sm.flags & kFlagFlip) NegateAcc<CFG>(acc);
    if (sm.flags & kFlagLast) {
      const GemmTile tile = p.tiles[sm.tile];
      const GemmGroup g = p.groups[tile.group];
      const uint32_t row0 = g.row_begin + uint32_t(tile.tm) * WBM, col0 = uint32_t(tile.tn) * WBN;
      bool write_c = true;
      if (tile.nsplit > 1) {
        // deterministic split-K: park this unit's partial tile, the last unit to arrive adds all of them
        double2 *slots = static_cast<double2 *>(p.partials) + (unsigned long long) tile.part_base * (WBM * WBN);
        double2 *mine = slots + (unsigned long long) tile.split * (WBM * WBN) + g4 * WBN + q * 8 + 2 * t4;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < NTMAX; ++j) {
            mine[i * 8 * WBN + j * 32] = AccValue<CFG>(acc, i, j, 0);
            mine[i * 8 * WBN + j * 32 + 1] = AccValue<CFG>(acc, i, j, 1);
          }
        __threadfence();
        ConsumerBarrier();
        if (warp == 0 && lane == 0) s_last = atomicAdd(&p.counters[2 + tile.ctr], 1u) == uint32_t(tile.nsplit) - 1u ? 1u : 0u;
        ConsumerBarrier();
        write_c = false;
        if (s_last != 0) FixupTile<CFG, MCAST, ACC>(p, tile, g, q, g4, t4);
      }
#ifdef QLB200_EXP_NOEPI
      if (acc[0][0][0][0] != 12345.678) write_c = false;
#endif
      if (write_c) {
        for (uint32_t d = 0; d < p.n_out; ++d) {     // n_out > 1: fused exchange, the same tile goes to every NVLink peer
          double2 *Cg = static_cast<double2 *>(p.c_out[d]) + g.c_off;
          if constexpr (ACC) {
            const double2 *Ci = static_cast<const double2 *>(p.c_in) + g.c_in_off;
            const double2 z = make_double2(0.0, 0.0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t row = row0 + i * 8 + g4;
              if (row >= g.row_end) continue;
              const unsigned long long at = (unsigned long long) row * g.n;
              if constexpr (NTMAX == 4)
                StoreAccRow(p, Cg + at, Ci + at, col0 + q * 8 + 2 * t4, g.n, g.beta_on, 4, AccValue<CFG>(acc, i, 0, 0), AccValue<CFG>(acc, i, 0, 1),
                            AccValue<CFG>(acc, i, 1, 0), AccValue<CFG>(acc, i, 1, 1), AccValue<CFG>(acc, i, 2, 0), AccValue<CFG>(acc, i, 2, 1),
                            AccValue<CFG>(acc, i, NTMAX - 1, 0), AccValue<CFG>(acc, i, NTMAX - 1, 1));
              else
                StoreAccRow(p, Cg + at, Ci + at, col0 + q * 8 + 2 * t4, g.n, g.beta_on, 3, AccValue<CFG>(acc, i, 0, 0), AccValue<CFG>(acc, i, 0, 1),
                            AccValue<CFG>(acc, i, 1, 0), AccValue<CFG>(acc, i, 1, 1), AccValue<CFG>(acc, i, 2, 0), AccValue<CFG>(acc, i, 2, 1), z, z);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t row = row0 + i * 8 + g4;
              if (row >= g.row_end) continue;
#pragma unroll
              for (int j = 0; j < NTMAX; ++j) {
                const uint32_t col = col0 + (q + 4 * j) * 8 + 2 * t4;
                double2 *dst = Cg + (unsigned long long) row * g.n + col;
                if (col < g.n) StoreOut(dst, AccValue<CFG>(acc, i, j, 0), MCAST);
                if (col + 1 < g.n) StoreOut(dst + 1, AccValue<CFG>(acc, i, j, 1), MCAST);
              }
            }
          }
        }
      }
    }
  }
}

template<class CFG>
cudaError_t Launch(const GemmParams &p, int num_sms, cudaStream_t stream) {
  constexpr size_t smem = WsSmem<CFG, CFG::STAGES>::kBytes;
  const uint32_t cap = 2u * uint32_t(num_sms);     // two resident CTAs per SM
  const uint32_t grid = p.seg != nullptr ? p.nseg : (p.ntiles < cap ? p.ntiles : cap);
  if (p.accum) GemmWsCplx<CFG, CFG::STAGES, false, true><<<grid, kWsThreads, smem, stream>>>(p);    // never with mcast (checked by the caller)
  else if (p.mcast) GemmWsCplx<CFG, CFG::STAGES, true, false><<<grid, kWsThreads, smem, stream>>>(p);
  else GemmWsCplx<CFG, CFG::STAGES, false, false><<<grid, kWsThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

template<class CFG>
cudaError_t Configure() {
  constexpr int smem = int(WsSmem<CFG, CFG::STAGES>::kBytes);
  cudaError_t e = cudaFuncSetAttribute(GemmWsCplx<CFG, CFG::STAGES, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(GemmWsCplx<CFG, CFG::STAGES, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(GemmWsCplx<CFG, CFG::STAGES, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

}  // namespace

cudaError_t ConfigureWsKernel() {
  cudaError_t e = Configure<Cfg4M>();
  if (e != cudaSuccess) return e;
  return Configure<Cfg3M>();
}

cudaError_t LaunchGemmWsCplx(const GemmParams &p, bool three_m, int num_sms, cudaStream_t stream) {
  if (p.ntiles == 0) return cudaSuccess;
  return three_m ? Launch<Cfg3M>(p, num_sms, stream) : Launch<Cfg4M>(p, num_sms, stream);
}

}  // namespace qlb200
