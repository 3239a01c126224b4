// ws_common.cuh -- pieces shared by the warp-specialised grouped GEMM kernels (gemm_ws.cu complex,
// gemm_ws_real.cu double): mbarrier / cp.async wrappers, the per-stage meta word, thread layout.
#ifndef QLB200_WS_COMMON_CUH
#define QLB200_WS_COMMON_CUH

#include "common.cuh"

namespace qlb200 {
namespace {

constexpr int kConsumerWarps = 4;
// consumer warpgroup + producer warpgroup.  Two CTAs x 8 warps leave 128
// registers per thread at launch; the budget is re-split at run time with setmaxnreg (consumers 208,
// producer warpgroup 48: per SM sub-partition 2 x (208 + 48) = 512 registers per lane).
constexpr int kWsThreads = (kConsumerWarps + 4) * 32;
constexpr uint32_t kFlagFirst = 1u, kFlagLast = 2u, kFlagNeg = 4u, kFlagATrans = 8u, kFlagBTrans = 16u, kFlagFlip = 32u;
constexpr uint32_t kSentinel = 0xffffffffu;
constexpr int kProducerWarps = 4;
constexpr uint32_t kFullArrivals = kProducerWarps * 32 + 1;   // async copy arrivals of every producer lane + the meta release

// named barrier 2: the consumer warps only (split-K fix-up)
__device__ __forceinline__ void ConsumerBarrier() { asm volatile("bar.sync 2, %0;" ::"n"(kConsumerWarps * 32) : "memory"); }
// named barrier 1: the producer warps only
__device__ __forceinline__ void ProducerBarrier() { asm volatile("bar.sync 1, %0;" ::"n"(kProducerWarps * 32) : "memory"); }

__device__ __forceinline__ uint32_t SmemAddr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void MbarInit(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void MbarArrive(uint64_t *bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(SmemAddr(bar)) : "memory");
}
// arrives on `bar` once all cp.async issued so far by this thread have landed (does not change the expected count)
__device__ __forceinline__ void CpAsyncMbarArrive(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(SmemAddr(bar)) : "memory");
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
__device__ __forceinline__ void CpAsync16Z(uint32_t smem, const void *gmem, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem), "l"(gmem), "r"(sz) : "memory");
}

__device__ __forceinline__ void DmmaNv(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// sign flip on the integer pipe (a DADD would compete with DMMA for the FP64 datapath)
__device__ __forceinline__ double FlipSign(double v, uint32_t mask) {
  return __hiloint2double(__double2hiint(v) ^ int(mask), __double2loint(v));
}

// L2 prefetch hint of the 8-byte copies: .L2::256B (SASS LDGSTS.E.LTC256B) measured +1 % on row-major A operands, neutral
// elsewhere (exp/r2_call22.sh); 0 = none, 128 = .L2::128B
#ifndef QLB200_CP8_L2PF
#define QLB200_CP8_L2PF 256
#endif
#if QLB200_CP8_L2PF == 256
#define QLB200_CP8_QUAL ".L2::256B"
#elif QLB200_CP8_L2PF == 128
#define QLB200_CP8_QUAL ".L2::128B"
#else
#define QLB200_CP8_QUAL ""
#endif
__device__ __forceinline__ void CpAsync8Z(uint32_t smem, const void *gmem, bool pred) {
  const int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global" QLB200_CP8_QUAL " [%0], [%1], 8, %2;" ::"r"(smem), "l"(gmem), "r"(sz) : "memory");
}

__device__ __forceinline__ void PrefetchL2(const void *gmem) { asm volatile("prefetch.global.L2 [%0];" ::"l"(gmem) : "memory"); }

// flags word of a stage: bits 0..5 first / last / neg / A, B transposed / flip the accumulators after the stage,
// 8..11 valid m8 groups, 16..20 valid n8 groups, 24..26 valid k4 steps
struct StageMeta { uint32_t tile, flags; };

}  // namespace
}  // namespace qlb200
#endif
