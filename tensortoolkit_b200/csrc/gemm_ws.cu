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
#include "common.cuh"
#include "ws_common.cuh"

namespace qlb200 {

namespace {

constexpr int WBM = kWsBM, WBN = kWsBN, WBK = kWsBK;
// Shared-memory tile layouts (units: complex elements = one 16-byte bank group); every fragment load
// of a quarter-warp (lanes g4 in {2p, 2p+1}, t4 in 0..3) hits 8 distinct bank groups:
//   A row-major   [32 m][12]      (12*g4 + t4)  mod 8 distinct
//   A transposed  [8 k][34]       (34*t4 + g4)  mod 8 = 2*t4 + g4 distinct
//   B row-major   [8 k][130]      (130*t4 + g4) mod 8 = 2*t4 + g4 distinct
//   B transposed  [128 n][8] with the k4 halves of odd rows swapped (k ^ 4*(n&1)): 4*(g4&1) + t4 distinct
constexpr int WLDA = WBK + 4, WLDAT = WBM + 2;
constexpr int WLDB = WBN + 2;
constexpr int A_ELEMS = WBM * WLDA, B_ELEMS = WBK * WLDB, STAGE_ELEMS = A_ELEMS + B_ELEMS;
static_assert(WBK * WLDAT <= A_ELEMS && WBN * WBK <= B_ELEMS, "transposed tiles must fit the stage");

template<int STAGES>
struct WsSmem {
  static constexpr size_t kBytes = size_t(STAGES) * STAGE_ELEMS * sizeof(double2) + 2 * STAGES * sizeof(uint64_t) +
                                   STAGES * sizeof(StageMeta);
};

// Fragment addressing of one stage (element units, relative to the stage's A / B regions).
struct FragAddr {
  const double2 *a, *b;     // lane's base inside the A / B tile
  uint32_t a_i, a_k1;       // A: + i * a_i (m8 group)  + ks * a_k1 (second k4 step)
  uint32_t b_j, b_k0, b_k1; // B: + j * b_j (owned n8 group) + b_k0 / b_k1 (first / second k4 step)
};

// One k-stage (WBK = 8 -> two k4 steps) of a warp's sub-tile: MT valid m8 row groups x NT valid n8 column
// groups.  Specialised at compile time so that skipped MMAs are not even issued.
template<int MT, int NT>
__device__ __forceinline__ void ComputeStage(double (&cr)[4][4][2], double (&ci)[4][4][2], const FragAddr &f, uint32_t smask) {
#pragma unroll
  for (int ks = 0; ks < WBK / 4; ++ks) {
    double ax[MT], ay[MT], nay[MT];
    double2 b[NT];
    const double2 *pa = f.a + (ks ? f.a_k1 : 0u);
    const double2 *pb = f.b + (ks ? f.b_k1 : f.b_k0);
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      const double2 a = pa[i * f.a_i];
      ax[i] = FlipSign(a.x, smask);
      ay[i] = FlipSign(a.y, smask);
      nay[i] = FlipSign(a.y, smask ^ 0x80000000u);
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) b[j] = pb[j * f.b_j];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        DmmaNv(cr[i][j][0], cr[i][j][1], ax[i], b[j].x);
        DmmaNv(ci[i][j][0], ci[i][j][1], ax[i], b[j].y);
      }
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        DmmaNv(cr[i][j][0], cr[i][j][1], nay[i], b[j].y);
        DmmaNv(ci[i][j][0], ci[i][j][1], ay[i], b[j].x);
      }
  }
}

template<int MT>
__device__ __forceinline__ void ComputeStageN(double (&cr)[4][4][2], double (&ci)[4][4][2], const FragAddr &f, uint32_t smask, int nt) {
  switch (nt) {
    case 4: ComputeStage<MT, 4>(cr, ci, f, smask); break;
    case 3: ComputeStage<MT, 3>(cr, ci, f, smask); break;
    case 2: ComputeStage<MT, 2>(cr, ci, f, smask); break;
    case 1: ComputeStage<MT, 1>(cr, ci, f, smask); break;
    default: break;
  }
}

// Split-K fix-up, run by the unit that arrived last: add the tile's partial tiles in slot order (a fixed
// order, whichever unit happens to be last) one m8 row group at a time and write C.  Kept out of line:
// it needs only a handful of registers and must not disturb the register allocation of the main loop.
__device__ __noinline__ void FixupTile(const GemmParams &p, const GemmTile &tile, const GemmGroup &g, int q, int g4, int t4) {
  __threadfence();
  const uint32_t row0 = g.row_begin + uint32_t(tile.tm) * WBM, col0 = uint32_t(tile.tn) * WBN;
  const double2 *src0 = static_cast<const double2 *>(p.partials) + (unsigned long long) tile.part_base * (WBM * WBN) +
                        g4 * WBN + q * 8 + 2 * t4;
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    double2 sum[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) sum[j][0] = sum[j][1] = make_double2(0.0, 0.0);
    const double2 *src = src0 + i * 8 * WBN;
    for (uint32_t sp = 0; sp < tile.nsplit; ++sp, src += WBM * WBN) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double2 v0 = __ldcg(src + j * 32), v1 = __ldcg(src + j * 32 + 1);
        sum[j][0].x += v0.x; sum[j][0].y += v0.y; sum[j][1].x += v1.x; sum[j][1].y += v1.y;
      }
    }
    const uint32_t row = row0 + i * 8 + g4;
    if (row < g.row_end) {
      for (uint32_t d = 0; d < p.n_out; ++d) {
        double2 *Cg = static_cast<double2 *>(p.c_out[d]) + g.c_off + (unsigned long long) row * g.n;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t col = col0 + (q + 4 * j) * 8 + 2 * t4;
          if (col < g.n) Cg[col] = sum[j][0];
          if (col + 1 < g.n) Cg[col + 1] = sum[j][1];
        }
      }
    }
  }
  if (q == 0 && g4 == 0 && t4 == 0) p.counters[2 + tile.ctr] = 0;
}

template<int STAGES>
__global__ void __launch_bounds__(kWsThreads, 2)
GemmWsCplx(const __grid_constant__ GemmParams p) {
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
    const uint32_t a_kc = lane & 7, a_r = lane >> 3;
    uint32_t it = 0, tcount = 0;
    for (;; ++tcount) {
      if (pw == 0 && lane == 0) s_tile[tcount & 1u] = atomicAdd(&p.counters[0], 1u);
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
        const uint32_t nst = (task.k + WBK - 1) / WBK;
        const uint32_t st_lo = max(sidx, tile.s_begin), st_hi = min(sidx + nst, tile.s_end);
        const uint32_t st_base = sidx;
        sidx += nst;
        if (st_lo >= st_hi) continue;
        const double2 *aBase = static_cast<const double2 *>((task.flags & kTaskASrc) ? p.a_src : p.a_ws) + task.a_off;
        const double2 *bBase = static_cast<const double2 *>((task.flags & kTaskBSrc) ? p.b_src : p.b_ws) + task.b_off;
        const bool ta = (task.flags & kTaskATrans) != 0, tb = (task.flags & kTaskBTrans) != 0;
        const uint32_t tflags = extents | (task.sign < 0 ? kFlagNeg : 0u) | (ta ? kFlagATrans : 0u) | (tb ? kFlagBTrans : 0u);
        for (uint32_t st = st_lo, k0 = (st_lo - st_base) * WBK; st < st_hi; ++st, ++it, k0 += WBK) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
          MbarWait(&empty[s], ph ^ 1u);
          const uint32_t sA = SmemAddr(stages + size_t(s) * STAGE_ELEMS);
          const uint32_t sB = sA + A_ELEMS * 16u;
          if (!ta) {   // A row-major m x k: a lane copies element (a_r + 4r, a_kc) of the 32 x 8 tile
            const uint32_t kk = k0 + a_kc;
            const bool kok = kk < task.k;
            const double2 *src = aBase + (unsigned long long) (row0 + a_r) * task.k + kk;
#pragma unroll
            for (uint32_t rr = 0; rr < 2; ++rr) {
              const uint32_t r = 2u * pw + rr, row = a_r + 4u * r;
              const bool ok = kok && row < rows;
              CpAsync16Z(sA + (row * WLDA + a_kc) * 16u, ok ? src + (unsigned long long) (4u * r) * task.k : aBase, ok);
            }
          } else {     // A stored k x m: 8 k-rows of 32 contiguous elements
            const double2 *src = aBase + (unsigned long long) k0 * g.m + row0 + lane;
            const bool mok = uint32_t(lane) < rows;
#pragma unroll
            for (uint32_t rr = 0; rr < 2; ++rr) {
              const uint32_t kr = 2u * pw + rr;
              const bool ok = mok && k0 + kr < task.k;
              CpAsync16Z(sA + (kr * WLDAT + lane) * 16u, ok ? src + (unsigned long long) kr * g.m : aBase, ok);
            }
          }
          if (!tb) {   // B row-major k x n: 8 k-rows x 128 columns, 512 contiguous bytes per copy
#pragma unroll
            for (uint32_t rr = 0; rr < 2; ++rr) {
              const uint32_t kr = 2u * pw + rr;
              const bool rok = k0 + kr < task.k;
              const double2 *src = bBase + (unsigned long long) (k0 + kr) * g.n + col0 + lane;
#pragma unroll
              for (uint32_t c = 0; c < 4; ++c) {
                const uint32_t col = lane + 32u * c;
                const bool ok = rok && col < cols;
                CpAsync16Z(sB + (kr * WLDB + col) * 16u, ok ? src + 32u * c : bBase, ok);
              }
            }
          } else {     // B stored n x k: a lane copies element (a_r + 4r, a_kc) of the 128 x 8 tile
            const uint32_t kk = k0 + a_kc;
            const bool kok = kk < task.k;
            const double2 *src = bBase + (unsigned long long) (col0 + a_r) * task.k + kk;
#pragma unroll
            for (uint32_t rr = 0; rr < 8; ++rr) {
              const uint32_t r = 8u * pw + rr, nl = a_r + 4u * r;
              const bool ok = kok && nl < cols;
              CpAsync16Z(sB + (nl * WBK + (a_kc ^ ((nl & 1u) << 2))) * 16u, ok ? src + (unsigned long long) (4u * r) * task.k : bBase, ok);
            }
          }
          CpAsyncMbarArrive(&full[s]);
          if (pw == 0 && lane == 0) {
            uint32_t fl = tflags;
            if (st == tile.s_begin) fl |= kFlagFirst;
            if (st + 1 == tile.s_end) fl |= kFlagLast;
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
  double cr[4][4][2], ci[4][4][2];
  uint32_t s = 0, ph = 0;     // ring position and phase parity
  for (;; ph ^= (++s == uint32_t(STAGES)) ? 1u : 0u, s = (s == uint32_t(STAGES)) ? 0u : s) {
    MbarWait(&full[s], ph);
    const StageMeta sm = meta[s];
    if (sm.tile == kSentinel) break;
    if (sm.flags & kFlagFirst) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;
    }
    const int mt = int((sm.flags >> 8) & 0xfu);
    const int n8 = int((sm.flags >> 16) & 0x1fu);
    const int nt = n8 > q ? (n8 - q + 3) >> 2 : 0;
    const double2 *tileA = stages + size_t(s) * STAGE_ELEMS;
    const double2 *tileB = tileA + A_ELEMS;
    FragAddr f;
    if (sm.flags & kFlagATrans) { f.a = tileA + t4 * WLDAT + g4; f.a_i = 8; f.a_k1 = 4 * WLDAT; }
    else { f.a = tileA + g4 * WLDA + t4; f.a_i = 8 * WLDA; f.a_k1 = 4; }
    if (sm.flags & kFlagBTrans) { f.b = tileB + (q * 8 + g4) * WBK + t4; f.b_j = 32 * WBK; f.b_k0 = (g4 & 1) << 2; f.b_k1 = f.b_k0 ^ 4u; }
    else { f.b = tileB + t4 * WLDB + q * 8 + g4; f.b_j = 32; f.b_k0 = 0; f.b_k1 = 4 * WLDB; }
    const uint32_t smask = (sm.flags & kFlagNeg) ? 0x80000000u : 0u;
    if (mt == 4 && nt == 4) {
      ComputeStage<4, 4>(cr, ci, f, smask);
    } else {
      switch (mt) {
        case 4: ComputeStageN<4>(cr, ci, f, smask, nt); break;
        case 3: ComputeStageN<3>(cr, ci, f, smask, nt); break;
        case 2: ComputeStageN<2>(cr, ci, f, smask, nt); break;
        default: ComputeStageN<1>(cr, ci, f, smask, nt); break;
      }
    }
    __syncwarp();
    if (lane == 0) MbarArrive(&empty[s]);
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
          for (int j = 0; j < 4; ++j) {
            mine[i * 8 * WBN + j * 32] = make_double2(cr[i][j][0], ci[i][j][0]);
            mine[i * 8 * WBN + j * 32 + 1] = make_double2(cr[i][j][1], ci[i][j][1]);
          }
        __threadfence();
        ConsumerBarrier();
        if (warp == 0 && lane == 0) s_last = atomicAdd(&p.counters[2 + tile.ctr], 1u) == uint32_t(tile.nsplit) - 1u ? 1u : 0u;
        ConsumerBarrier();
        write_c = false;
        if (s_last != 0) FixupTile(p, tile, g, q, g4, t4);
      }
      if (write_c) {
        for (uint32_t d = 0; d < p.n_out; ++d) {     // n_out > 1: fused exchange, the same tile goes to every NVLink peer
          double2 *Cg = static_cast<double2 *>(p.c_out[d]) + g.c_off;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t row = row0 + i * 8 + g4;
            if (row >= g.row_end) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t col = col0 + (q + 4 * j) * 8 + 2 * t4;
              double2 *dst = Cg + (unsigned long long) row * g.n + col;
              if (col < g.n) dst[0] = make_double2(cr[i][j][0], ci[i][j][0]);
              if (col + 1 < g.n) dst[1] = make_double2(cr[i][j][1], ci[i][j][1]);
            }
          }
        }
      }
    }
  }
}

constexpr int kWsStages = 5;

}  // namespace

cudaError_t ConfigureWsKernel() {
  return cudaFuncSetAttribute(GemmWsCplx<kWsStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(WsSmem<kWsStages>::kBytes));
}

cudaError_t LaunchGemmWsCplx(const GemmParams &p, int num_sms, cudaStream_t stream) {
  if (p.ntiles == 0) return cudaSuccess;
  constexpr size_t smem = WsSmem<kWsStages>::kBytes;
  const uint32_t cap = 2u * uint32_t(num_sms);     // two resident CTAs per SM
  const uint32_t grid = p.ntiles < cap ? p.ntiles : cap;
  GemmWsCplx<kWsStages><<<grid, kWsThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace qlb200
