// permute.cu -- batched N-D permute: every block of both operands in ONE launch.
//
// Replaces the per-block hp_numeric::TensorTranspose calls of the reference's contraction loop
// (include/qlten/qltensor/blk_spar_data_ten/global_operations.h:922-964; HPTT
// framework/hp_numeric/ten_trans.h:94-114, cuTENSOR :245-349) and the per-block loop of
// BlockSparseDataTensor::Transpose (raw_data_operations.h:201-218).
// Semantics (ten_trans.h:130-186): out[i_perm[0], i_perm[1], ...] = scale * in[i_0, i_1, ...], both
// row-major, i.e. output axis j is input axis perm[j].  Pure data movement: bit-exact.
//
// HBM-bound.  Persistent CTAs walk a flat tile list (prefix sums over blocks, binary search), stage
// a TI x TO tile in shared memory so that global reads run along the source-fastest axis and
// global writes along the destination-fastest axis.  Algorithmic bytes: 2 * elements * sizeof(T).
#include "common.cuh"

namespace qlb200 {

namespace {

constexpr int kPermThreads = 256;
constexpr int kPermSmemElems = 2304;   // (TI|1) * TO never exceeds this (host tiler guarantees)

template<typename T> __device__ __forceinline__ T ScaleBy(T v, float s);
template<> __device__ __forceinline__ double ScaleBy<double>(double v, float s) { return s < 0.f ? -v : v; }
template<> __device__ __forceinline__ double2 ScaleBy<double2>(double2 v, float s) {
  return s < 0.f ? make_double2(-v.x, -v.y) : v;
}

template<typename T>
__global__ void __launch_bounds__(kPermThreads)
PermuteKernel(const PermBlk *__restrict__ blks, const uint32_t *__restrict__ tile_base, uint32_t nblk,
              uint32_t ntiles, const T *__restrict__ srcA, const T *__restrict__ srcB,
              T *__restrict__ dstA, T *__restrict__ dstB, uint32_t allow_bulk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *s = reinterpret_cast<T *>(smem_raw);
  __shared__ PermBlk sd;
  __shared__ uint32_t s_blk;
  __shared__ uint16_t s_off[kPermSmemElems];   // run mode: shared-memory offset of every element of a destination piece
  __shared__ __align__(8) uint64_t s_bar;      // bulk path: completion barrier of the global -> shared copies
  const int tid = threadIdx.x;
  uint32_t cur_blk = 0xffffffffu;
  uint32_t bulk_phase = 0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&s_bar))) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // --- locate the block owning this tile (largest b with tile_base[b] <= tile) ---
    if (tid == 0) {
      uint32_t lo = 0, hi = nblk;
      while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (tile_base[mid] <= tile) lo = mid; else hi = mid;
      }
      s_blk = lo;
    }
    __syncthreads();
    const uint32_t b = s_blk;
    if (b != cur_blk) {
      // cooperative copy of the descriptor into shared memory
      const uint32_t *g = reinterpret_cast<const uint32_t *>(blks + b);
      uint32_t *d = reinterpret_cast<uint32_t *>(&sd);
      for (int i = tid; i < int(sizeof(PermBlk) / 4); i += kPermThreads) d[i] = g[i];
      cur_blk = b;
    }
    __syncthreads();

    uint32_t lt = tile - tile_base[b];
    const uint32_t nd = sd.nd, jin = sd.jin, jout = sd.jout, V = sd.vec;
    const uint32_t ti0 = (lt % sd.nti) * sd.TI; lt /= sd.nti;
    const uint32_t to0 = (lt % sd.nto) * sd.TO; lt /= sd.nto;
    // sstr[jin] is 1 except in run mode, where jin is the axis behind the run in the source (stride V)
    unsigned long long so = sd.src_off + (unsigned long long) ti0 * sd.sstr[jin], dofs = sd.dst_off + (unsigned long long) ti0 * sd.dstr[jin];
    if (jout != jin) {
      so += (unsigned long long) to0 * sd.sstr[jout];
      dofs += (unsigned long long) to0 * sd.dstr[jout];
    }
    for (int j = int(nd) - 1; j >= 0; --j) {
      if (uint32_t(j) == jin || uint32_t(j) == jout || (V != 0u && uint32_t(j) == nd - 1)) continue;
      const uint32_t e = sd.ext[j];
      const uint32_t c = lt % e; lt /= e;
      so += (unsigned long long) c * sd.sstr[j];
      dofs += (unsigned long long) c * sd.dstr[j];
    }
    const uint32_t TIa = min(sd.TI, sd.ext[jin] - ti0);
    const uint32_t TOa = (jout != jin) ? min(sd.TO, sd.ext[jout] - to0) : 1u;
    const T *__restrict__ src = (sd.src_sel ? srcB : srcA) + so;
    T *__restrict__ dst = (sd.src_sel ? dstB : dstA) + dofs;
    const float scale = sd.scale;
    const uint32_t s_out = sd.sstr[jout], d_in = sd.dstr[jin], d_out = sd.dstr[jout];

    if (V != 0u && sd.bulk != 0u && allow_bulk != 0u) {
      // run mode, bulk asynchronous copies (no per-element instructions at all): thread 0 queues TOa copies of TIa*V contiguous
      // source elements into shared memory, completing on the mbarrier; then every thread queues one shared -> global copy
      // per V-element run of the tile (TIa*TOa runs), commits them and waits until shared memory has been read
      const uint32_t L = TIa * V, P = sd.TI * V;
      const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&s_bar));
      const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(s));
      if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(uint32_t(TOa * L * sizeof(T))) : "memory");
        for (uint32_t to = 0; to < TOa; ++to)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sbase + uint32_t(to * P * sizeof(T))),
                       "l"(src + (unsigned long long) to * s_out), "r"(uint32_t(L * sizeof(T))), "r"(bar)
                       : "memory");
      }
      {
        uint32_t done = 0;
        while (!done)
          asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(bulk_phase) : "memory");
        bulk_phase ^= 1u;
      }
      const uint32_t nruns = TIa * TOa;
      for (uint32_t idx = tid; idx < nruns; idx += kPermThreads) {
        const uint32_t ti = idx / TOa, to = idx - ti * TOa;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (unsigned long long) ti * d_in + (unsigned long long) to * V),
                     "r"(sbase + uint32_t((to * P + ti * V) * sizeof(T))), "r"(uint32_t(V * sizeof(T)))
                     : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    } else if (V != 0u) {
      // run mode: TOa pieces of TIa*V contiguous source elements in, TIa pieces of TOa*V contiguous destination elements out
      const uint32_t L = TIa * V, P = sd.TI * V + 1u, M = TOa * V;
      // index arithmetic without a division per element: each thread advances its (piece, offset) pair by the block
      // size; the (to, v) decomposition of a destination offset comes from a small shared table
      for (uint32_t y = tid; y < M; y += kPermThreads) { const uint32_t to = y / V; s_off[y] = uint16_t(to * P + (y - to * V)); }
      {
        const uint32_t q = kPermThreads / L, r = kPermThreads - q * L, total = TOa * L;
        uint32_t to = tid / L, x = tid - to * L;
        for (uint32_t idx = tid; idx < total; idx += kPermThreads) {
          s[to * P + x] = src[(unsigned long long) to * s_out + x];
          to += q; x += r;
          if (x >= L) { x -= L; ++to; }
        }
      }
      __syncthreads();
      {
        const uint32_t q = kPermThreads / M, r = kPermThreads - q * M, total = TIa * M;
        uint32_t ti = tid / M, y = tid - ti * M;
        for (uint32_t idx = tid; idx < total; idx += kPermThreads) {
          dst[(unsigned long long) ti * d_in + y] = ScaleBy(s[s_off[y] + ti * V], scale);
          ti += q; y += r;
          if (y >= M) { y -= M; ++ti; }
        }
      }
    } else if (jin == nd - 1) {
      // source-fastest axis is also destination-fastest: contiguous runs, no staging needed
      const uint32_t tx = tid & ((1u << sd.txi_log2) - 1u), ty = tid >> sd.txi_log2;
      const uint32_t TX = 1u << sd.txi_log2, RY = kPermThreads >> sd.txi_log2;
      for (uint32_t to = ty; to < TOa; to += RY) {
        const T *sp = src + (unsigned long long) to * s_out;
        T *dp = dst + (unsigned long long) to * d_out;
#pragma unroll 4
        for (uint32_t ti = tx; ti < TIa; ti += TX) dp[ti] = ScaleBy(sp[ti], scale);
      }
    } else {
      const uint32_t TIp = sd.TI | 1u;
      {
        const uint32_t tx = tid & ((1u << sd.txi_log2) - 1u), ty = tid >> sd.txi_log2;
        const uint32_t TX = 1u << sd.txi_log2, RY = kPermThreads >> sd.txi_log2;
        for (uint32_t to = ty; to < TOa; to += RY) {
          const T *sp = src + (unsigned long long) to * s_out;
          T *ss = s + to * TIp;
#pragma unroll 4
          for (uint32_t ti = tx; ti < TIa; ti += TX) ss[ti] = sp[ti];
        }
      }
      __syncthreads();
      {
        const uint32_t tx = tid & ((1u << sd.txo_log2) - 1u), ty = tid >> sd.txo_log2;
        const uint32_t TX = 1u << sd.txo_log2, RY = kPermThreads >> sd.txo_log2;
        for (uint32_t ti = ty; ti < TIa; ti += RY) {
          T *dp = dst + (unsigned long long) ti * d_in;
          const T *ss = s + ti;
#pragma unroll 4
          for (uint32_t to = tx; to < TOa; to += TX) dp[to] = ScaleBy(ss[to * TIp], scale);
        }
      }
    }
    __syncthreads();   // shared tile / descriptor are reused by the next iteration
  }
}

// dst[dst_off + j] = beta * src[src_off + j] over a short list of ranges (accumulate form: output blocks the contraction
// does not touch).  Every CTA walks all ranges and takes a grid-strided share of each: coalesced, HBM-bound, no tables.
template<typename T>
__global__ void __launch_bounds__(256)
ScaleCopyRanges(const unsigned long long *__restrict__ r3, uint32_t nranges, const T *__restrict__ src, T *__restrict__ dst,
                double beta_re, double beta_im) {
  const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;
  for (uint32_t r = 0; r < nranges; ++r) {
    const unsigned long long so = r3[3 * r], dof = r3[3 * r + 1], len = r3[3 * r + 2];
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
      if constexpr (sizeof(T) == 16) {
        if (beta_re == 0.0 && beta_im == 0.0) { dst[dof + i] = make_double2(0.0, 0.0); continue; }     // exact zeros, nothing read
        const double2 v = src[so + i];
        dst[dof + i] = make_double2(beta_re * v.x - beta_im * v.y, beta_re * v.y + beta_im * v.x);
      } else {
        if (beta_re == 0.0) { dst[dof + i] = 0.0; continue; }
        dst[dof + i] = beta_re * src[so + i];
      }
    }
  }
}

// Fan-out copy: src[0 .. n16) (16-byte units) -> the same range of every destination.  MCAST: one destination, an NVSwitch
// multicast mapping (multimem.st: the store leaves the GPU once, the switch writes every replica).
template<bool MCAST>
__global__ void __launch_bounds__(256)
FanOutCopy(const uint4 *__restrict__ src, unsigned long long n16, uint4 *d0, uint4 *d1, uint4 *d2, uint4 *d3, uint4 *d4, uint4 *d5,
           uint4 *d6, uint4 *d7, uint32_t ndst) {
  const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    const uint4 v = src[i];
    if constexpr (MCAST) {
      asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d0 + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    } else {
      d0[i] = v;
      if (ndst > 1) d1[i] = v;
      if (ndst > 2) d2[i] = v;
      if (ndst > 3) d3[i] = v;
      if (ndst > 4) d4[i] = v;
      if (ndst > 5) d5[i] = v;
      if (ndst > 6) d6[i] = v;
      if (ndst > 7) d7[i] = v;
    }
  }
}

}  // namespace

cudaError_t LaunchFanOutCopy(const void *src, unsigned long long bytes, void *const *dst, uint32_t ndst, bool mcast, int num_sms,
                             cudaStream_t stream) {
  if (bytes == 0) return cudaSuccess;
  const unsigned long long n16 = bytes / 16;
  uint4 *d[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (uint32_t i = 0; i < ndst && i < 8; ++i) d[i] = static_cast<uint4 *>(dst[i]);
  const unsigned long long want = (n16 + 255) / 256;
  const uint32_t grid = uint32_t(want < (unsigned long long) num_sms * 4 ? want : (unsigned long long) num_sms * 4);
  if (mcast) FanOutCopy<true><<<grid, 256, 0, stream>>>(static_cast<const uint4 *>(src), n16, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], 1);
  else FanOutCopy<false><<<grid, 256, 0, stream>>>(static_cast<const uint4 *>(src), n16, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], ndst);
  return cudaGetLastError();
}

cudaError_t LaunchScaleCopyRanges(int dtype, const unsigned long long *ranges3, uint32_t nranges, const void *src, void *dst,
                                  double beta_re, double beta_im, int num_sms, cudaStream_t stream) {
  if (nranges == 0) return cudaSuccess;
  const uint32_t grid = uint32_t(num_sms) * 8u;
  if (dtype == 0)
    ScaleCopyRanges<double><<<grid, 256, 0, stream>>>(ranges3, nranges, static_cast<const double *>(src), static_cast<double *>(dst), beta_re, beta_im);
  else
    ScaleCopyRanges<double2><<<grid, 256, 0, stream>>>(ranges3, nranges, static_cast<const double2 *>(src), static_cast<double2 *>(dst), beta_re, beta_im);
  return cudaGetLastError();
}

cudaError_t LaunchPermute(int dtype, const PermBlk *blks, const uint32_t *tile_base, uint32_t nblk,
                          uint32_t ntiles, const void *srcA, const void *srcB, void *dstA, void *dstB,
                          int num_sms, cudaStream_t stream) {
  if (ntiles == 0) return cudaSuccess;
  const uint32_t grid = ntiles < uint32_t(num_sms) * 6u ? ntiles : uint32_t(num_sms) * 6u;
  // the plan marks runs whose ELEMENT offsets are 16-byte aligned; bulk copies also need 16-byte aligned base pointers, which
  // only the caller's buffers can break (a double buffer at an odd element of a larger allocation)
  const auto misaligned = [](const void *q) { return q != nullptr && (reinterpret_cast<uintptr_t>(q) & 15u) != 0; };
  const uint32_t allow_bulk = (misaligned(srcA) || misaligned(srcB) || misaligned(dstA) || misaligned(dstB)) ? 0u : 1u;
  if (dtype == 0) {
    const size_t smem = kPermSmemElems * sizeof(double);
    PermuteKernel<double><<<grid, kPermThreads, smem, stream>>>(
        blks, tile_base, nblk, ntiles, static_cast<const double *>(srcA), static_cast<const double *>(srcB),
        static_cast<double *>(dstA), static_cast<double *>(dstB), allow_bulk);
  } else {
    const size_t smem = kPermSmemElems * sizeof(double2);
    PermuteKernel<double2><<<grid, kPermThreads, smem, stream>>>(
        blks, tile_base, nblk, ntiles, static_cast<const double2 *>(srcA), static_cast<const double2 *>(srcB),
        static_cast<double2 *>(dstA), static_cast<double2 *>(dstB), allow_bulk);
  }
  return cudaGetLastError();
}

}  // namespace qlb200
