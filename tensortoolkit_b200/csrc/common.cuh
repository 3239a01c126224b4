// common.cuh -- device descriptor tables shared by the plan builder and the kernels.
#ifndef QLB200_COMMON_CUH
#define QLB200_COMMON_CUH

#include <cuda_runtime.h>
#include <cstdint>
#include <string>

namespace qlb200 {

// ------------------------------------------------------------------------------------------------
// Batched permute.  One PermBlk per block, already canonicalised on the host:
//   * size-1 axes dropped, axes that stay adjacent (and in order) merged;
//   * axes listed in OUTPUT order (ext[nd-1] is the output-fastest axis), sstr[] = source stride.
// The tile of a block spans TI elements of the source-fastest axis (`jin`, source stride 1) and TO
// elements of the output-fastest axis (nd-1); all remaining axes enumerate tiles.  When the two
// axes coincide (jin == nd-1) the block is a batch of contiguous runs ("row copy") and TO
// enumerates the next-outer output axis instead -- unless the runs are short, see PermBlk::vec.
// ------------------------------------------------------------------------------------------------
constexpr int kPermMaxDims = 8;
constexpr int kPermTileElems = 2048;   // elements staged in shared memory per tile

struct PermBlk {
  unsigned long long src_off, dst_off;  // elements, relative to the source / destination buffer
  uint32_t ext[kPermMaxDims];           // extents in output order
  uint32_t sstr[kPermMaxDims];          // source stride of each output axis
  uint32_t dstr[kPermMaxDims];          // destination stride of each output axis (row-major over ext)
  uint32_t nd;
  uint32_t jin;                         // output axis with source stride 1
  uint32_t jout;                        // axis tiled by TO (nd-1, or the next-outer one for row copies)
  uint32_t TI, TO;                      // tile extents
  uint32_t nti, nto;                    // tiles along jin / jout
  uint32_t txi_log2, txo_log2;          // log2 of the thread-row width used in the load / store phase
  uint32_t src_sel;                     // 0 = operand A buffer, 1 = operand B buffer
  float scale;                          // +1 / -1 (fermionic whole-tensor transpose), applied on the fly
  uint32_t bulk;                        // run mode only: every run is 16-byte aligned on both sides and needs no scaling, so the
                                        // tile moves as bulk asynchronous copies (cp.async.bulk, SASS UBLKCP): TO pieces
                                        // global -> shared completing on an mbarrier, then one shared -> global copy per run
  uint32_t vec;                         // 0, or V = ext[nd-1]: "run" mode -- the fastest axis is the same on both sides but
                                        // short (V elements); the tile is then a 2-D transposition of V-element runs:
                                        // jin = source-next-fastest axis (source stride V, tiled by TI), jout = nd-2
                                        // (destination stride V, tiled by TO); loads read TI*V, stores write TO*V contiguous
};

// ------------------------------------------------------------------------------------------------
// Grouped GEMM.  A "group" is one output block; its tasks are the matched input pairs that
// accumulate into it.  Offsets are in elements.
// ------------------------------------------------------------------------------------------------
struct GemmTask {
  unsigned long long a_off, b_off;  // elements, into the caller's buffer (kTask?Src) or the permuted workspace
  uint32_t k;
  int16_t sign;                     // +1 / -1
  uint16_t flags;                   // kTask* bits
  // B as a k x n matrix VIEW of the stored block (not for kTaskBTrans): element (kk, col) lies at
  //     b_off + kk * b_rs + (col / b_run) * b_cs + col % b_run
  // Row-major k x n (also every permuted copy): b_rs = b_run = n, b_cs = 0.  A block stored (n1, k, n2) and contracted
  // over k -- permutation {1, 0, 2}, the ragged stress test's B -- is read in place with b_rs = b_run = n2, b_cs = k * n2:
  // rows of the view are n1 runs of n2 contiguous elements, so the producer's copies stay contiguous per run.
  uint32_t b_rs, b_cs, b_run;
  uint32_t pad_;
};
constexpr uint16_t kTaskASrc = 1;    // A block is read from the caller's buffer, not the workspace
constexpr uint16_t kTaskATrans = 2;  // ... where it is stored as a row-major k x m matrix
constexpr uint16_t kTaskBSrc = 4;    // B block is read from the caller's buffer
constexpr uint16_t kTaskBTrans = 8;  // ... where it is stored as a row-major n x k matrix

struct GemmGroup {
  unsigned long long c_off;
  uint32_t m, n;
  uint32_t task_begin, task_end;
  uint32_t row_begin, row_end;      // rows of the block this plan computes (multi-GPU slabs)
  // accumulate plans (C_out = beta * C_in + alpha * sum of pairs): where the block lies in the EXISTING output, and whether
  // it is read at all (0 for blocks the contraction adds to the output topology: their first-task beta is zero)
  unsigned long long c_in_off;
  uint32_t beta_on;
  uint32_t pad_;
};

// DMMA work unit: rows [tm*BM, ..) x cols [tn*BN, ..) of a group, k-stages [s_begin, s_end) of the
// group's concatenated k loop.  A tile whose k loop is long compared with a CTA's share of the launch
// is split into `nsplit` units (deterministic split-K): every unit stores its partial tile into slot
// part_base + split of the workspace; the unit that arrives last (counter `ctr`) adds the partials in
// slot order 0..nsplit-1 -- a fixed order, whichever unit happens to be last -- and writes C.
struct GemmTile {
  uint32_t group;
  uint16_t tm, tn;
  uint32_t s_begin, s_end;
  uint32_t part_base, ctr;
  uint16_t split, nsplit;
  uint32_t pad_;
};

struct SkinnyItem {     // rows [row0, row0 + skinny_sub * (kSkinnyElems / n)) of a narrow group
  uint32_t group;
  uint32_t row0;
};

// ------------------------------------------------------------------------------------------------
// Matrix-free axis operations (axis_kernel.cu).  A block is viewed as (P0, X1, P1, X2, P2): X1 / X2 the two target axes.
// ------------------------------------------------------------------------------------------------
constexpr int kAxisMaxDim = 8;    // largest operator block edge the kernel keeps in its shared-memory coefficient rows
struct AxisOut {                  // one output block
  unsigned long long out_off;
  uint32_t size, J1, P1, J2, P2;  // P0 = size / (J1 P1 J2 P2)
  uint32_t term_begin, term_end;  // flat terms
  uint32_t pad_;
};
struct AxisFlatTerm {             // one (input block, i1, i2) slice feeding an output block
  unsigned long long in_base;     // input offset of element (p0 = 0, i1, p1 = 0, i2, p2 = 0)
  unsigned long long c1_off, c2_off;   // row i1 of the op1 block / row i2 of the op2 block (J1 / J2 consecutive values)
  uint32_t sp0, sp1;              // input strides of p0 and p1 (they depend on the INPUT block's own I1, I2)
};
struct AxisItem { uint32_t out, elem0; };   // a 1024-element chunk of an output block
cudaError_t LaunchAxisApply(int dtype, const AxisOut *outs, const AxisFlatTerm *terms, const AxisItem *items, uint32_t nitems, const void *in,
                            const void *op1, const void *op2, void *out, int num_sms, cudaStream_t stream);

constexpr int kMaxOut = 8;        // output replicas (this GPU + NVLink peers)

struct GemmParams {
  const void *a_src, *a_ws, *b_src, *b_ws;   // caller's raw buffers / permuted workspace (set per launch)
  void *c_out[kMaxOut];                      // every output element is stored to c_out[0 .. n_out): the caller's C and,
  uint32_t n_out;                            // for a fused multi-GPU exchange, the same buffer on each NVLink peer
  uint32_t mcast;                            // c_out[0] is an NVSwitch multicast address: one multimem.st reaches every GPU
  const GemmTask *tasks;
  const GemmGroup *groups;
  const GemmTile *tiles;
  const SkinnyItem *items;
  uint32_t ntiles, nitems;
  uint32_t skinny_sub;              // sub-chunks of kSkinnyElems outputs per narrow-pair work item
  const uint32_t *seg;              // stream-K: CTA b runs units [seg[b], seg[b+1]) (nullptr: units are pulled from counters[0])
  uint32_t nseg;
  unsigned int *counters;           // [0] next tile, [1] finished CTAs, [2 + ctr] split-K arrivals (all self-resetting)
  void *partials;                   // split-K partial tiles, slot = BM x BN elements
  // accumulate form (qlb200_execute_accum): C_out = beta * C_in + alpha * (sum of the block's pairs); alpha / beta are
  // (re, im), im ignored for real tensors.  accum == 0: plain contraction, C_in is never read.
  uint32_t accum;
  const void *c_in;
  double alpha_re, alpha_im, beta_re, beta_im;
};

#ifdef __CUDACC__
// Epilogue of the accumulate form: v <- alpha * v + (beta_on ? beta * C_in[idx] : 0).
__device__ __forceinline__ double AxpbyOut(const GemmParams &p, double v, const double *cin, bool beta_on) {
  double r = p.alpha_re * v;
  if (beta_on) r = fma(p.beta_re, *cin, r);
  return r;
}
__device__ __forceinline__ double2 AxpbyOut(const GemmParams &p, double2 v, const double2 *cin, bool beta_on) {
  double2 r = make_double2(p.alpha_re * v.x - p.alpha_im * v.y, p.alpha_re * v.y + p.alpha_im * v.x);
  if (beta_on) {
    const double2 c = *cin;
    r.x += p.beta_re * c.x - p.beta_im * c.y;
    r.y += p.beta_re * c.y + p.beta_im * c.x;
  }
  return r;
}

// Output stores.  With `mcast` the address is a multicast mapping of the result buffer (NVLS): the store leaves the
// GPU once and the NVSwitch replicates it into every GPU's copy -- multimem.st is the only legal access to such memory.
__device__ __forceinline__ void StoreOut(double2 *dst, double2 v, uint32_t mcast) {
  if (mcast) {   // 16 bytes as four 32-bit lanes (multimem.st has no .v2.f64 form); SASS: one STG.E.128
    asm volatile(
        "{\n .reg .b32 q0, q1, q2, q3;\n mov.b64 {q0, q1}, %1;\n mov.b64 {q2, q3}, %2;\n"
        " multimem.st.weak.global.v4.f32 [%0], {q0, q1, q2, q3};\n}" ::"l"(dst), "d"(v.x), "d"(v.y)
        : "memory");
  } else {
    *dst = v;
  }
}
__device__ __forceinline__ void StoreOut(double *dst, double v, uint32_t mcast) {
  if (mcast) asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(dst), "d"(v) : "memory");
  else *dst = v;
}
#endif

inline std::string CudaErr(const char *what, cudaError_t e) {
  return std::string(what) + ": " + cudaGetErrorString(e);
}

// launchers (defined in permute.cu / gemm.cu); all enqueue on `stream` and return the launch error
cudaError_t LaunchPermute(int dtype, const PermBlk *blks, const uint32_t *tile_base, uint32_t nblk,
                          uint32_t ntiles, const void *srcA, const void *srcB, void *dstA, void *dstB,
                          int num_sms, cudaStream_t stream);
cudaError_t LaunchGemmSkinny(int dtype, const GemmParams &p, int num_sms, cudaStream_t stream);
// dst[dst_off[i] + j] = beta * src[src_off[i] + j], j < len[i]: the output blocks of an accumulate call that the contraction
// does not touch (ScaleUntouchedOutputBlocks_ / ExpandOutputTopology_, contract_contiguous_axes.h:567-671).  Ranges are
// {src_off, dst_off, len} triples in elements.
cudaError_t LaunchScaleCopyRanges(int dtype, const unsigned long long *ranges3, uint32_t nranges, const void *src, void *dst,
                                  double beta_re, double beta_im, int num_sms, cudaStream_t stream);
// src[0 .. bytes) -> the same bytes of every dst[i] (unicast peer pointers), or of dst[0] = an NVSwitch multicast mapping;
// 16-byte granularity (bytes and all pointers multiples of 16)
cudaError_t LaunchFanOutCopy(const void *src, unsigned long long bytes, void *const *dst, uint32_t ndst, bool mcast, int num_sms,
                             cudaStream_t stream);
cudaError_t ConfigureKernels();   // one-time cudaFuncSetAttribute calls
// warp-specialised complex kernel (gemm_ws.cu), CTA tile kWsBM x kWsBN (4M arithmetic) or kWsBM x kWs3mBN (3M)
cudaError_t LaunchGemmWsCplx(const GemmParams &p, bool three_m, int num_sms, cudaStream_t stream);
cudaError_t ConfigureWsKernel();
// warp-specialised real-double kernel (gemm_ws_real.cu), CTA tile kWsRealBM x kWsRealBN
cudaError_t LaunchGemmWsReal(const GemmParams &p, int num_sms, cudaStream_t stream);
cudaError_t ConfigureWsRealKernel();

// tile shapes of the DMMA kernel, needed by the host-side tiler
constexpr int kWsBM = 32, kWsBN = 128, kWs3mBN = 96;       // warp-specialised complex kernel (4M / 3M tile width)
constexpr int kWsRealBM = 64, kWsRealBN = 128;             // warp-specialised real kernel
constexpr int kWsBK = 8, kWsRealBK = 16;                   // k extent of one pipeline stage
constexpr int kWs3mBK = 16;                                // 3M kernel: two half-stages share one barrier round trip
constexpr int kWs3mStages = kWs3mBK == 16 ? 3 : 5;
// narrow-pair kernel: 4 outputs per thread keep it at 64 registers -> 4 CTAs (32 warps) per SM; the kernel is
// latency-bound, so resident warps (loads in flight) matter more than per-thread reuse (measured: 8 per thread /
// 2 CTAs 0.83 ms, 4 / 4 CTAs 0.69 ms, 2 / 6 CTAs 0.99 ms for the two MPO steps of the D=4096 apply; a one-thread-per-row
// variant -- every A value loaded once instead of n times, n strided stores per thread, 2 CTAs -- measured 0.90 ms)
constexpr int kSkinnyMaxN = 8, kSkinnyMaxK = 32, kSkinnyThreads = 256, kSkinnyPerThread = 4, kSkinnyMinCtas = 4;
constexpr int kSkinnyElems = kSkinnyThreads * kSkinnyPerThread;   // output elements per sub-chunk of a work item
// sub-chunks per work item: the descriptor fetch (a chain of dependent global loads) and the term table are amortised over
// them.  Chosen per plan (PlanHost::skinny_sub): as many as possible, up to 8, while every resident CTA still gets about
// eight items (measured on the MPO steps, HBM fraction with 1 / 4 / 8 sub-chunks: D=4096 complex 0.70 / 0.81 / 0.79,
// Hubbard D=8192 0.62 / 0.72 / 0.78, D=1024 0.30 / 0.32 / 0.26)
constexpr int kSkinnyMaxSub = 8;

}  // namespace qlb200
#endif
