// gemm_ws.cu -- warp-specialised grouped complex-FP64 GEMM (the headline kernel).
//
// Same contract as gemm.cu (C = sum over the group's pairs of sign * A[m x k] * B[k x n], written
// once, deterministic), different machine mapping:
//
//   * 1 producer warp + 8 consumer warps per CTA, one persistent CTA per SM.
//   * The producer moves operand tiles global -> shared with the bulk asynchronous copy engine
//     (cp.async.bulk, SASS UBLKCP; one copy per tile row, completion counted in bytes on an
//     mbarrier).  Complex elements are 16 bytes, so every row start / length meets the engine's
//     16-byte rule without any constraint on block offsets or leading dimensions.
//   * Stages are handed over with full/empty mbarriers only -- there is no CTA-wide barrier in the
//     main loop, so the FP64 tensor pipe (DMMA.8x8x4) never waits for the slowest warp.
//   * Every stage carries {tile id, first/last/sign flags}; the producer runs ahead across tile
//     boundaries (dynamic atomic tile counter), so the next tile's prologue overlaps this tile's
//     epilogue stores.
//   * Rows beyond m / columns beyond n are simply not copied (they only feed outputs that are
//     never stored); only the K tail of a pair is zero-filled, in both operands.
#include "common.cuh"

namespace qlb200 {

namespace {

constexpr int kConsumerWarps = 8;
// 2 consumer warpgroups + 1 producer warpgroup (only its first warp works).  With 3 warps per SM
// sub-partition the register file allows 168 registers per thread; the consumers need ~230 for a
// 32x32 complex accumulator tile, so the register budget is re-split at run time with setmaxnreg
// (producer warpgroup 40, consumer warpgroups 232: 2*232 + 40 = 504 <= 512 per lane and sub-partition).
constexpr int kWsThreads = (kConsumerWarps + 4) * 32;
constexpr int WBK = 8;                 // k extent of one stage
constexpr int WLDA = WBK + 4;          // 12 complex per A row: (row*12 + k) mod 8 distinct over a quarter-warp
constexpr uint32_t kFlagFirst = 1u, kFlagLast = 2u, kFlagNeg = 4u;
constexpr uint32_t kSentinel = 0xffffffffu;

__device__ __forceinline__ uint32_t SmemAddr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void MbarInit(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void MbarArrive(uint64_t *bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(SmemAddr(bar)) : "memory");
}
__device__ __forceinline__ void MbarArriveExpectTx(uint64_t *bar, uint32_t bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(SmemAddr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void MbarWait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "WAIT_LOOP:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra WAIT_DONE;\n"
      " bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}" ::"r"(SmemAddr(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void BulkCopyG2S(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(SmemAddr(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(SmemAddr(bar))
               : "memory");
}

__device__ __forceinline__ void DmmaNv(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct StageMeta { uint32_t tile, flags; };

template<int MT, int NT, int STAGES>
struct WsCfg {
  static constexpr int BM = 16 * MT;            // 2 warp rows x MT m8-tiles
  static constexpr int BN = 32 * NT;            // 4 warp cols x NT n8-tiles
  static constexpr int LDB = BN + 2;            // (k*LDB + n) mod 8 = 2k + n distinct over a quarter-warp
  static constexpr int A_ELEMS = BM * WLDA;
  static constexpr int B_ELEMS = WBK * LDB;
  static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
  static constexpr size_t SMEM = size_t(STAGES) * STAGE_ELEMS * sizeof(double2) + 2 * STAGES * sizeof(uint64_t) +
                                 STAGES * sizeof(StageMeta) + 64;
};

template<int MT, int NT, int STAGES>
__global__ void __launch_bounds__(kWsThreads, 1)
GemmWsCplx(GemmParams p, const double2 *__restrict__ A, const double2 *__restrict__ B, double2 *__restrict__ C) {
  using Cfg = WsCfg<MT, NT, STAGES>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2 *stages = reinterpret_cast<double2 *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + size_t(STAGES) * Cfg::STAGE_ELEMS * sizeof(double2));
  uint64_t *empty = full + STAGES;
  StageMeta *meta = reinterpret_cast<StageMeta *>(empty + STAGES);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { MbarInit(&full[s], 1); MbarInit(&empty[s], kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp >= kConsumerWarps) {
    // ============================== producer warpgroup ==============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp != kConsumerWarps) return;
    uint32_t it = 0;
    for (;;) {
      uint32_t tile_id = 0;
      if (lane == 0) tile_id = atomicAdd(&p.counters[0], 1u);
      tile_id = __shfl_sync(0xffffffffu, tile_id, 0);
      if (tile_id >= p.ntiles) break;
      const GemmTile tile = p.tiles[tile_id];
      const GemmGroup g = p.groups[tile.group];
      const uint32_t row0 = g.row_begin + uint32_t(tile.tm) * Cfg::BM, col0 = uint32_t(tile.tn) * Cfg::BN;
      const uint32_t rows = min(uint32_t(Cfg::BM), g.row_end - row0), cols = min(uint32_t(Cfg::BN), g.n - col0);
      for (uint32_t t = g.task_begin; t < g.task_end; ++t) {
        const GemmTask task = p.tasks[t];
        const double2 *gA = A + task.a_off + (unsigned long long) row0 * task.k;
        const double2 *gB = B + task.b_off + col0;
        for (uint32_t k0 = 0; k0 < task.k; k0 += WBK, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
          MbarWait(&empty[s], ph ^ 1u);
          double2 *sA = stages + size_t(s) * Cfg::STAGE_ELEMS;
          double2 *sB = sA + Cfg::A_ELEMS;
          const uint32_t kt = min(uint32_t(WBK), task.k - k0);
          if (kt < uint32_t(WBK)) {     // K tail: exact zeros in both operands
            const uint32_t zc = WBK - kt;
            for (uint32_t i = lane; i < uint32_t(Cfg::BM) * zc; i += 32) sA[(i / zc) * WLDA + kt + (i % zc)] = make_double2(0.0, 0.0);
            for (uint32_t i = lane; i < zc * uint32_t(Cfg::BN); i += 32) sB[(kt + i / Cfg::BN) * Cfg::LDB + (i % Cfg::BN)] = make_double2(0.0, 0.0);
          }
          __syncwarp();
          if (lane == 0) {
            uint32_t fl = (task.sign < 0 ? kFlagNeg : 0u);
            if (t == g.task_begin && k0 == 0) fl |= kFlagFirst;
            if (t + 1 == g.task_end && k0 + WBK >= task.k) fl |= kFlagLast;
            meta[s].tile = tile_id; meta[s].flags = fl;
            MbarArriveExpectTx(&full[s], (rows * kt + kt * cols) * 16u);
          }
          __syncwarp();
          for (uint32_t r = lane; r < rows; r += 32)
            BulkCopyG2S(sA + r * WLDA, gA + (unsigned long long) r * task.k + k0, kt * 16u, &full[s]);
          if (lane < kt)
            BulkCopyG2S(sB + lane * Cfg::LDB, gB + (unsigned long long) (k0 + lane) * g.n, cols * 16u, &full[s]);
        }
      }
    }
    // sentinel stage: tells every consumer warp to leave
    {
      const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
      MbarWait(&empty[s], ph ^ 1u);
      if (lane == 0) { meta[s].tile = kSentinel; meta[s].flags = 0; MbarArrive(&full[s]); }
    }
    if (lane == 0) {
      __threadfence();
      if (atomicAdd(&p.counters[1], 1u) == gridDim.x - 1) { p.counters[0] = 0; p.counters[1] = 0; __threadfence(); }
    }
    return;
  }

  // ================================ consumer warps ================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int wm0 = (warp >> 2) * (8 * MT), wn0 = (warp & 3) * (8 * NT);
  const int g4 = lane >> 2, t4 = lane & 3;
  double cr[MT][NT][2], ci[MT][NT][2];
  uint32_t it = 0;
  for (;; ++it) {
    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
    MbarWait(&full[s], ph);
    const StageMeta sm = meta[s];
    if (sm.tile == kSentinel) break;
    if (sm.flags & kFlagFirst) {
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;
    }
    const double2 *cA = stages + size_t(s) * Cfg::STAGE_ELEMS + (wm0 + g4) * WLDA + t4;
    const double2 *cB = stages + size_t(s) * Cfg::STAGE_ELEMS + Cfg::A_ELEMS + t4 * Cfg::LDB + wn0 + g4;
    const bool neg = (sm.flags & kFlagNeg) != 0;
#pragma unroll
    for (int ks = 0; ks < WBK / 4; ++ks) {
      double2 a[MT], b[NT];
      double nai[MT];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        a[i] = cA[i * 8 * WLDA + ks * 4];
        if (neg) { a[i].x = -a[i].x; a[i].y = -a[i].y; }
        nai[i] = -a[i].y;
      }
#pragma unroll
      for (int j = 0; j < NT; ++j) b[j] = cB[ks * 4 * Cfg::LDB + j * 8];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          DmmaNv(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
          DmmaNv(ci[i][j][0], ci[i][j][1], a[i].x, b[j].y);
        }
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          DmmaNv(cr[i][j][0], cr[i][j][1], nai[i], b[j].y);
          DmmaNv(ci[i][j][0], ci[i][j][1], a[i].y, b[j].x);
        }
    }
    __syncwarp();
    if (lane == 0) MbarArrive(&empty[s]);
    if (sm.flags & kFlagLast) {
      const GemmTile tile = p.tiles[sm.tile];
      const GemmGroup g = p.groups[tile.group];
      const uint32_t row0 = g.row_begin + uint32_t(tile.tm) * Cfg::BM, col0 = uint32_t(tile.tn) * Cfg::BN;
      double2 *Cg = C + g.c_off;
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const uint32_t row = row0 + wm0 + i * 8 + g4;
        if (row >= g.row_end) continue;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const uint32_t col = col0 + wn0 + j * 8 + 2 * t4;
          double2 *dst = Cg + (unsigned long long) row * g.n + col;
          if (col < g.n) dst[0] = make_double2(cr[i][j][0], ci[i][j][0]);
          if (col + 1 < g.n) dst[1] = make_double2(cr[i][j][1], ci[i][j][1]);
        }
      }
    }
  }
}

template<int MT, int NT, int STAGES>
cudaError_t LaunchOne(const GemmParams &p, const void *A, const void *B, void *C, int num_sms, cudaStream_t stream) {
  using Cfg = WsCfg<MT, NT, STAGES>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(GemmWsCplx<MT, NT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const uint32_t grid = p.ntiles < uint32_t(num_sms) ? p.ntiles : uint32_t(num_sms);
  GemmWsCplx<MT, NT, STAGES><<<grid, kWsThreads, Cfg::SMEM, stream>>>(p, static_cast<const double2 *>(A), static_cast<const double2 *>(B),
                                                                     static_cast<double2 *>(C));
  return cudaGetLastError();
}

}  // namespace

// shape: 0 = 64x128, 1 = 32x128, 2 = 64x64, 3 = 32x64  (CTA tile rows x cols)
cudaError_t LaunchGemmWsCplx(int shape, const GemmParams &p, const void *A, const void *B, void *C, int num_sms,
                             cudaStream_t stream) {
  if (p.ntiles == 0) return cudaSuccess;
  switch (shape) {
    case 0: return LaunchOne<4, 4, 6>(p, A, B, C, num_sms, stream);
    case 1: return LaunchOne<2, 4, 6>(p, A, B, C, num_sms, stream);
    case 2: return LaunchOne<4, 2, 6>(p, A, B, C, num_sms, stream);
    case 3: return LaunchOne<2, 2, 6>(p, A, B, C, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

void WsTileShape(int shape, int *bm, int *bn) {
  static const int kBm[4] = {64, 32, 64, 32}, kBn[4] = {128, 128, 64, 64};
  *bm = kBm[shape & 3]; *bn = kBn[shape & 3];
}

}  // namespace qlb200
