// axis_kernel.cu -- matrix-free application of one or two small rank-2 operators to tensor axes, axis order preserved:
//     out[p0, j1, p1, j2, p2] = sum over contributing input blocks, i1, i2 of  in[p0, i1, p1, i2, p2] * op1[i1, j1] * op2[i2, j2]
// (qlten::dmrg::ApplyRank2ToAxisPreserveOrder / ApplyTwoRank2ToAxesPreserveOrder, tensor_manipulation/dmrg/axis_ops.h:2889-3125;
// per-block kernels AddRank2AxisBlock :1695-1775 and AddTwoRank2AxesBlockGemm :1932-1986).  Every block is viewed as a
// rank-5 array (the axes in front of, between and behind the two target axes merged); with one operator j2 = i2 = 1.
//
// HBM-bound: each output element is written once; each input element is read once per output block it feeds (the
// reference's per-block GEMMs re-read the output instead, beta = 1).  One launch for the whole tensor: persistent CTAs walk
// a list of 1024-element output chunks; the (block triple, i1, i2) terms of a chunk's output block are flattened into a
// shared-memory table {input slice base, its two outer strides, op1 row, op2 row}; a thread owns four output elements
// (consecutive threads = consecutive addresses along the innermost axis, so loads and stores are coalesced).
#include "common.cuh"

namespace qlb200 {

namespace {

constexpr int kAxisThreads = 256, kAxisPerThread = 4, kAxisChunk = kAxisThreads * kAxisPerThread;
constexpr int kAxisTermChunk = 32;

template<typename T> __device__ __forceinline__ T Mul(T a, T b);
template<> __device__ __forceinline__ double Mul<double>(double a, double b) { return a * b; }
template<> __device__ __forceinline__ double2 Mul<double2>(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template<typename T> __device__ __forceinline__ void Fma(T &acc, T a, T b);
template<> __device__ __forceinline__ void Fma<double>(double &acc, double a, double b) { acc = fma(a, b, acc); }
template<> __device__ __forceinline__ void Fma<double2>(double2 &acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
template<typename T> __device__ __forceinline__ T Zero();
template<> __device__ __forceinline__ double Zero<double>() { return 0.0; }
template<> __device__ __forceinline__ double2 Zero<double2>() { return make_double2(0.0, 0.0); }
template<typename T> __device__ __forceinline__ T One();
template<> __device__ __forceinline__ double One<double>() { return 1.0; }
template<> __device__ __forceinline__ double2 One<double2>() { return make_double2(1.0, 0.0); }

template<typename T>
__global__ void __launch_bounds__(kAxisThreads, 4)
AxisApply(const AxisOut *__restrict__ outs, const AxisFlatTerm *__restrict__ terms, const AxisItem *__restrict__ items, uint32_t nitems,
          const T *__restrict__ in, const T *__restrict__ op1, const T *__restrict__ op2, T *__restrict__ out) {
  __shared__ AxisFlatTerm s_term[kAxisTermChunk];
  __shared__ T s_c1[kAxisTermChunk][kAxisMaxDim], s_c2[kAxisTermChunk][kAxisMaxDim];
  const uint32_t tid = threadIdx.x;
  for (uint32_t it = blockIdx.x; it < nitems; it += gridDim.x) {
    const AxisItem item = items[it];
    const AxisOut o = outs[item.out];
    const uint32_t total = min(uint32_t(kAxisChunk), o.size - item.elem0);
    // decompose this thread's output elements once: (p0, j1, p1, j2, p2)
    uint32_t p0[kAxisPerThread], p1[kAxisPerThread], p2[kAxisPerThread], j1[kAxisPerThread], j2[kAxisPerThread];
    T acc[kAxisPerThread];
#pragma unroll
    for (int u = 0; u < kAxisPerThread; ++u) {
      uint32_t e = item.elem0 + tid + u * kAxisThreads;
      p2[u] = e % o.P2; e /= o.P2;
      j2[u] = e % o.J2; e /= o.J2;
      p1[u] = e % o.P1; e /= o.P1;
      j1[u] = e % o.J1; p0[u] = e / o.J1;
      acc[u] = Zero<T>();
    }
    for (uint32_t t0 = o.term_begin; t0 < o.term_end; t0 += kAxisTermChunk) {
      const uint32_t nt = min(uint32_t(kAxisTermChunk), o.term_end - t0);
      __syncthreads();            // previous chunk / item fully consumed
      if (tid < nt) s_term[tid] = terms[t0 + tid];
      for (uint32_t x = tid; x < nt * kAxisMaxDim; x += kAxisThreads) {
        const uint32_t t = x / kAxisMaxDim, j = x % kAxisMaxDim;
        const AxisFlatTerm ft = terms[t0 + t];
        s_c1[t][j] = j < o.J1 ? op1[ft.c1_off + j] : Zero<T>();
        s_c2[t][j] = (op2 != nullptr && j < o.J2) ? op2[ft.c2_off + j] : One<T>();
      }
      __syncthreads();
#pragma unroll 1
      for (uint32_t t = 0; t < nt; ++t) {
        const AxisFlatTerm ft = s_term[t];
        T v[kAxisPerThread];
#pragma unroll
        for (int u = 0; u < kAxisPerThread; ++u)
          if (tid + u * kAxisThreads < total) v[u] = in[ft.in_base + (unsigned long long) p0[u] * ft.sp0 + (unsigned long long) p1[u] * ft.sp1 + p2[u]];
#pragma unroll
        for (int u = 0; u < kAxisPerThread; ++u)
          if (tid + u * kAxisThreads < total) Fma(acc[u], v[u], Mul(s_c1[t][j1[u]], s_c2[t][j2[u]]));
      }
    }
    T *dst = out + o.out_off + item.elem0;
#pragma unroll
    for (int u = 0; u < kAxisPerThread; ++u)
      if (tid + u * kAxisThreads < total) dst[tid + u * kAxisThreads] = acc[u];
  }
}

}  // namespace

cudaError_t LaunchAxisApply(int dtype, const AxisOut *outs, const AxisFlatTerm *terms, const AxisItem *items, uint32_t nitems, const void *in,
                            const void *op1, const void *op2, void *out, int num_sms, cudaStream_t stream) {
  if (nitems == 0) return cudaSuccess;
  const uint32_t cap = uint32_t(num_sms) * 8u;
  const uint32_t grid = nitems < cap ? nitems : cap;
  if (dtype == 0)
    AxisApply<double><<<grid, kAxisThreads, 0, stream>>>(outs, terms, items, nitems, static_cast<const double *>(in), static_cast<const double *>(op1),
                                                         static_cast<const double *>(op2), static_cast<double *>(out));
  else
    AxisApply<double2><<<grid, kAxisThreads, 0, stream>>>(outs, terms, items, nitems, static_cast<const double2 *>(in), static_cast<const double2 *>(op1),
                                                          static_cast<const double2 *>(op2), static_cast<double2 *>(out));
  return cudaGetLastError();
}

}  // namespace qlb200
