// capi.cu -- the extern "C" surface declared in include/qlb200.h.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "matcher.h"
#include "plan.h"
#include "qlb200.h"

using namespace qlb200;

namespace {

thread_local std::string g_err;

int Fail(int code, const std::string &msg) { g_err = msg; return code; }

#define QL_CUDA(call)                                                      \
  do {                                                                     \
    cudaError_t e_ = (call);                                               \
    if (e_ != cudaSuccess) return Fail(QLB200_ERR_CUDA, CudaErr(#call, e_)); \
  } while (0)

template<typename T>
int Upload(const std::vector<T> &v, T **dst, cudaStream_t s) {
  *dst = nullptr;
  if (v.empty()) return QLB200_OK;
  QL_CUDA(cudaMalloc(reinterpret_cast<void **>(dst), v.size() * sizeof(T)));
  QL_CUDA(cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  return QLB200_OK;
}

// Grow-only arenas.  A captured CUDA graph bakes arena addresses into its kernel nodes, so an arena that a live graph
// may reference is never freed when a later plan needs more room: it is retired (kept allocated until the context's last
// graph is destroyed, or the context itself) and a new, larger one takes its place.  Growth is impossible while the
// stream is capturing (cudaMalloc / synchronise are illegal there): the caller must run the sequence once eagerly first.
int EnsureArena(qlb200_ctx *ctx, void **p, size_t *cap, size_t need) {
  if (need <= *cap) return QLB200_OK;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(ctx->stream, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone)
    return Fail(QLB200_ERR_UNSUPPORTED, "workspace arena must grow while the stream is being captured: run the sequence once before qlb200_graph_begin");
  QL_CUDA(cudaStreamSynchronize(ctx->stream));
  if (*p) {
    if (ctx->graphs_alive > 0) ctx->retired.push_back(*p);
    else QL_CUDA(cudaFree(*p));
    *p = nullptr; *cap = 0;
  }
  size_t want = need + need / 8 + 256;
  QL_CUDA(cudaMalloc(p, want));
  *cap = want;
  return QLB200_OK;
}

size_t ElemSize(int dtype) { return dtype == QLB200_C64 ? 16 : 8; }
size_t Align256(size_t v) { return (v + 255) & ~size_t(255); }

int UploadGemmTables(qlb200_plan *p) {
  cudaStream_t s = p->ctx->stream;
  int rc;
  if (p->d.tasks == nullptr) {
    if ((rc = Upload(p->h.tasks, &p->d.tasks, s)) != QLB200_OK) return rc;
  }
  if (p->d.groups) { cudaFree(p->d.groups); p->d.groups = nullptr; }
  if (p->d.tiles) { cudaFree(p->d.tiles); p->d.tiles = nullptr; }
  if (p->d.items) { cudaFree(p->d.items); p->d.items = nullptr; }
  if (p->d.seg) { cudaFree(p->d.seg); p->d.seg = nullptr; }
  if ((rc = Upload(p->h.seg, &p->d.seg, s)) != QLB200_OK) return rc;
  if ((rc = Upload(p->h.part_groups, &p->d.groups, s)) != QLB200_OK) return rc;
  if ((rc = Upload(p->h.tiles, &p->d.tiles, s)) != QLB200_OK) return rc;
  if ((rc = Upload(p->h.items, &p->d.items, s)) != QLB200_OK) return rc;
  // [0] next unit, [1] finished CTAs, [2..] split-K arrival counters; the kernels leave them all zero
  if (p->d.counters) { cudaFree(p->d.counters); p->d.counters = nullptr; }
  const size_t nctr = 2 + size_t(p->h.n_split_ctrs);
  QL_CUDA(cudaMalloc(reinterpret_cast<void **>(&p->d.counters), nctr * sizeof(unsigned int)));
  QL_CUDA(cudaMemsetAsync(p->d.counters, 0, nctr * sizeof(unsigned int), s));
  QL_CUDA(cudaStreamSynchronize(s));   // host vectors may change after return
  return QLB200_OK;
}

int FinishPlan(qlb200_ctx *ctx, qlb200_plan *p) {
  cudaStream_t s = ctx->stream;
  int rc;
  QL_CUDA(cudaSetDevice(ctx->device));
  if ((rc = Upload(p->h.perm_blks, &p->d.perm_blks, s)) != QLB200_OK) return rc;
  if ((rc = Upload(p->h.perm_tile_base, &p->d.perm_tile_base, s)) != QLB200_OK) return rc;
  return UploadGemmTables(p);
}

GemmParams MakeParams(const qlb200_plan *p, const void *A, const void *B, const void *wsA, const void *wsB, void *partials,
                      void *const *c_out, uint32_t n_out, uint32_t mcast = 0) {
  GemmParams gp;
  gp.a_src = A; gp.b_src = B; gp.a_ws = wsA; gp.b_ws = wsB; gp.partials = partials;
  for (uint32_t d = 0; d < uint32_t(kMaxOut); ++d) gp.c_out[d] = d < n_out ? c_out[d] : nullptr;
  gp.n_out = n_out;
  gp.mcast = mcast;
  gp.tasks = p->d.tasks; gp.groups = p->d.groups; gp.tiles = p->d.tiles; gp.items = p->d.items;
  gp.ntiles = static_cast<uint32_t>(p->h.tiles.size());
  gp.nitems = static_cast<uint32_t>(p->h.items.size());
  gp.skinny_sub = p->h.skinny_sub;
  gp.seg = p->h.seg.empty() ? nullptr : p->d.seg;
  gp.nseg = p->h.seg.empty() ? 0u : static_cast<uint32_t>(p->h.seg.size() - 1);
  gp.counters = p->d.counters;
  gp.accum = 0; gp.c_in = nullptr;
  gp.alpha_re = 1.0; gp.alpha_im = 0.0; gp.beta_re = 0.0; gp.beta_im = 0.0;
  return gp;
}

void SetPlanKnobs(const qlb200_ctx *ctx, uint32_t *flags, PlanHost *h) {
  h->num_sms = ctx ? ctx->num_sms : 148;
  // process-wide choice of the complex product: QLB200_COMPLEX_PRODUCT=4m plans every complex contraction with the
  // four-product kernel (componentwise-accurate small parts; see DESIGN.md section 4 "Why 3M, and when not")
  if (const char *cp = std::getenv("QLB200_COMPLEX_PRODUCT"))
    if (cp[0] == '4') *flags |= QLB200_PLAN_CPLX_4M;
}

size_t WsBytes(const qlb200_plan *p) {
  const size_t es = ElemSize(p->h.dtype);
  return Align256(p->h.ws_a_elems * es) + Align256(p->h.ws_b_elems * es) + p->partials_shift +
         Align256(p->h.n_part_slots * p->h.part_slot_elems * es);
}

}  // namespace

namespace qlb200 {
int FailWith(int code, const std::string &msg) { return Fail(code, msg); }   // for the other translation units (comm.cu)
}

void qlb200::DeviceTables::Free() {
  cudaFree(perm_blks); cudaFree(perm_tile_base); cudaFree(tasks); cudaFree(groups); cudaFree(tiles);
  cudaFree(items); cudaFree(counters); cudaFree(seg);
  perm_blks = nullptr; perm_tile_base = nullptr; tasks = nullptr; groups = nullptr; tiles = nullptr;
  items = nullptr; counters = nullptr; seg = nullptr;
}

extern "C" {

const char *qlb200_version(void) { return "qlb200 0.1 (sm_100a)"; }
const char *qlb200_last_error(void) { return g_err.c_str(); }

// ---- matcher -----------------------------------------------------------------------------------
struct qlb200_match { Match m; };

static int MatchCreate(const qlb200_shell *a, const qlb200_shell *b, int32_t nctrct, const int32_t *a_axes,
                       const int32_t *b_axes, int sel_axis, uint32_t sel_sector, qlb200_match **out) {
  if (!a || !b || !out || (nctrct > 0 && (!a_axes || !b_axes))) return Fail(QLB200_ERR_ARG, "null argument");
  qlb200_match *m = new (std::nothrow) qlb200_match();
  if (!m) return Fail(QLB200_ERR_NOMEM, "out of memory");
  std::string err = BuildMatch(a, b, nctrct, a_axes, b_axes, sel_axis, sel_sector, &m->m);
  if (!err.empty()) { delete m; return Fail(QLB200_ERR_ARG, err); }
  *out = m;
  return QLB200_OK;
}

int qlb200_match_create(const qlb200_shell *a, const qlb200_shell *b, int32_t nctrct, const int32_t *a_axes,
                        const int32_t *b_axes, qlb200_match **out) {
  return MatchCreate(a, b, nctrct, a_axes, b_axes, -1, 0, out);
}
int qlb200_match_create_1sector(const qlb200_shell *a, int32_t axis, uint32_t sector, const qlb200_shell *b,
                                int32_t nctrct, const int32_t *a_axes, const int32_t *b_axes, qlb200_match **out) {
  if (axis < 0) return Fail(QLB200_ERR_ARG, "negative 1-sector axis");
  return MatchCreate(a, b, nctrct, a_axes, b_axes, axis, sector, out);
}
int qlb200_match_create_contiguous(const qlb200_shell *a, const qlb200_shell *b, int32_t a_start, int32_t b_start,
                                   int32_t size, qlb200_match **out) {
  if (!a || !b || !out) return Fail(QLB200_ERR_ARG, "null argument");
  const int32_t ra = a->rank, rb = b->rank;
  if (ra < 1 || rb < 1 || ra > QLB200_MAX_RANK || rb > QLB200_MAX_RANK) return Fail(QLB200_ERR_ARG, "rank out of range");
  if (size < 0 || size > ra || size > rb || a_start < 0 || a_start >= ra || b_start < 0 || b_start >= rb)
    return Fail(QLB200_ERR_ARG, "bad contiguous axis range");
  int32_t a_axes[QLB200_MAX_RANK], b_axes[QLB200_MAX_RANK], a_saved[QLB200_MAX_RANK], b_saved[QLB200_MAX_RANK];
  for (int32_t i = 0; i < size; ++i) { a_axes[i] = (a_start + i) % ra; b_axes[i] = (b_start + i) % rb; }
  const int32_t a_end = (a_start + size) % ra, b_end = (b_start + size) % rb;
  for (int32_t i = 0; i < ra - size; ++i) a_saved[i] = (a_end + i) % ra;
  for (int32_t i = 0; i < rb - size; ++i) b_saved[i] = (b_end + i) % rb;
  qlb200_match *m = new (std::nothrow) qlb200_match();
  if (!m) return Fail(QLB200_ERR_NOMEM, "out of memory");
  std::string err = BuildMatch(a, b, size, a_axes, b_axes, -1, 0, &m->m, a_saved, b_saved, true);
  if (!err.empty()) { delete m; return Fail(QLB200_ERR_ARG, err); }
  *out = m;
  return QLB200_OK;
}
int32_t qlb200_match_saved_axes(const qlb200_match *m, int which, int32_t *axes_out) {
  if (!m) return 0;
  const std::vector<int> &v = which == 0 ? m->m.a_saved : m->m.b_saved;
  if (axes_out) for (size_t i = 0; i < v.size(); ++i) axes_out[i] = v[i];
  return static_cast<int32_t>(v.size());
}
void qlb200_match_destroy(qlb200_match *m) { delete m; }
int32_t qlb200_match_c_rank(const qlb200_match *m) { return m ? m->m.c_rank : 0; }
uint64_t qlb200_match_c_nblk(const qlb200_match *m) { return m ? m->m.c_blocks.size() : 0; }
uint64_t qlb200_match_c_elems(const qlb200_match *m) { return m ? m->m.c_elems : 0; }
uint64_t qlb200_match_ntask(const qlb200_match *m) { return m ? m->m.tasks.size() : 0; }
int qlb200_match_is_scalar(const qlb200_match *m) { return m && m->m.scalar ? 1 : 0; }
int qlb200_match_perm(const qlb200_match *m, int which, int32_t *perm_out) {
  if (!m || !perm_out) return Fail(QLB200_ERR_ARG, "null argument");
  const std::vector<int> &p = which == 0 ? m->m.a_perm : m->m.b_perm;
  for (size_t i = 0; i < p.size(); ++i) perm_out[i] = p[i];
  return which == 0 ? (m->m.a_need_trans ? 1 : 0) : (m->m.b_need_trans ? 1 : 0);
}
int qlb200_match_c_blocks(const qlb200_match *m, uint64_t *blk_idx, uint32_t *blk_coors, uint32_t *shape,
                          uint64_t *offset) {
  if (!m) return Fail(QLB200_ERR_ARG, "null argument");
  const int r = m->m.c_rank;
  for (size_t b = 0; b < m->m.c_blocks.size(); ++b) {
    const CBlock &cb = m->m.c_blocks[b];
    if (blk_idx) blk_idx[b] = cb.blk_idx;
    if (offset) offset[b] = cb.offset;
    for (int i = 0; i < r; ++i) {
      if (blk_coors) blk_coors[b * r + i] = cb.coors[i];
      if (shape) shape[b * r + i] = cb.shape[i];
    }
  }
  return QLB200_OK;
}
int qlb200_match_tasks(const qlb200_match *m, int order, qlb200_task *tasks_out) {
  if (!m || (!tasks_out && !m->m.tasks.empty())) return Fail(QLB200_ERR_ARG, "null argument");
  if (order == 0) {
    std::memcpy(tasks_out, m->m.tasks.data(), m->m.tasks.size() * sizeof(qlb200_task));
  } else {
    auto s = m->m.SortedTasks();
    std::memcpy(tasks_out, s.data(), s.size() * sizeof(qlb200_task));
  }
  return QLB200_OK;
}
int qlb200_estimate_cost(const qlb200_match *m, int dtype, qlb200_cost *out) {
  if (!m || !out) return Fail(QLB200_ERR_ARG, "null argument");
  EstimateCost(m->m, dtype, out);
  return QLB200_OK;
}

// ---- context -----------------------------------------------------------------------------------
int qlb200_ctx_create(int device, qlb200_ctx **out) {
  if (!out) return Fail(QLB200_ERR_ARG, "null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return Fail(QLB200_ERR_CUDA, e != cudaSuccess ? CudaErr("cudaGetDeviceCount", e) : "no CUDA device");
  if (device < 0 || device >= ndev) return Fail(QLB200_ERR_ARG, "device ordinal out of range");
  QL_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  QL_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return Fail(QLB200_ERR_UNSUPPORTED, "qlb200 kernels are built for sm_100a only; found sm_" + std::to_string(prop.major * 10 + prop.minor));
  qlb200_ctx *c = new (std::nothrow) qlb200_ctx();
  if (!c) return Fail(QLB200_ERR_NOMEM, "out of memory");
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete c; return Fail(QLB200_ERR_CUDA, CudaErr("cudaStreamCreate", e)); }
  e = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->alt, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
  if (e != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; return Fail(QLB200_ERR_CUDA, CudaErr("side stream / events", e)); }
  e = ConfigureKernels();
  if (e != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; return Fail(QLB200_ERR_CUDA, CudaErr("cudaFuncSetAttribute", e)); }
  *out = c;
  return QLB200_OK;
}
void qlb200_ctx_destroy(qlb200_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->side) cudaStreamSynchronize(ctx->side);
  cudaFree(ctx->ws); cudaFree(ctx->stage);
  for (void *r : ctx->retired) cudaFree(r);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  if (ctx->copy) { cudaStreamSynchronize(ctx->copy); cudaStreamDestroy(ctx->copy); }
  if (ctx->alt) { cudaStreamSynchronize(ctx->alt); cudaStreamDestroy(ctx->alt); }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  delete ctx;
}
int qlb200_ctx_sync(qlb200_ctx *ctx) { QL_CUDA(cudaStreamSynchronize(ctx->stream)); return QLB200_OK; }
void *qlb200_ctx_stream(qlb200_ctx *ctx) { return ctx->stream; }
int qlb200_ctx_set_stream(qlb200_ctx *ctx, void *cuda_stream) {
  QL_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  ctx->own_stream = false;
  return QLB200_OK;
}
uint64_t qlb200_ctx_launch_count(const qlb200_ctx *ctx) { return ctx->launches; }

int qlb200_dev_alloc(qlb200_ctx *ctx, size_t bytes, void **out) {
  QL_CUDA(cudaSetDevice(ctx->device));
  QL_CUDA(cudaMalloc(out, bytes ? bytes : 1));
  return QLB200_OK;
}
int qlb200_dev_free(qlb200_ctx *ctx, void *p) {
  QL_CUDA(cudaSetDevice(ctx->device));
  QL_CUDA(cudaFree(p));
  return QLB200_OK;
}
int qlb200_memcpy_h2d(qlb200_ctx *ctx, void *dst, const void *src, size_t bytes) {
  QL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return QLB200_OK;
}
int qlb200_memcpy_d2h(qlb200_ctx *ctx, void *dst, const void *src, size_t bytes) {
  QL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return QLB200_OK;
}
int qlb200_host_register(void *p, size_t bytes) { QL_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterDefault)); return QLB200_OK; }
int qlb200_host_unregister(void *p) { QL_CUDA(cudaHostUnregister(p)); return QLB200_OK; }

// ---- plans -------------------------------------------------------------------------------------
int qlb200_plan_create(qlb200_ctx *ctx, const qlb200_match *m, const qlb200_shell *, const qlb200_shell *,
                       int dtype, uint32_t flags, qlb200_plan **out) {
  if (!m || !out) return Fail(QLB200_ERR_ARG, "null argument");
  if (dtype != QLB200_F64 && dtype != QLB200_C64) return Fail(QLB200_ERR_ARG, "bad dtype");
  const Match &mm = m->m;
  qlb200_plan *p = new (std::nothrow) qlb200_plan();
  if (!p) return Fail(QLB200_ERR_NOMEM, "out of memory");
  p->ctx = ctx;
  SetPlanKnobs(ctx, &flags, &p->h);
  std::vector<int32_t> ap(mm.a_perm.begin(), mm.a_perm.end()), bp(mm.b_perm.begin(), mm.b_perm.end());
  std::string err = BuildPlanHost(dtype, flags, static_cast<int>(mm.a_ctrct.size()), mm.a.rank, ap.data(), mm.a.nblk,
                                  mm.a.shape.data(), mm.a.offset.data(), mm.a.elems, mm.b.rank, bp.data(), mm.b.nblk,
                                  mm.b.shape.data(), mm.b.offset.data(), mm.b.elems, mm.SortedTasks(), mm.c_elems, &p->h);
  if (!err.empty()) { delete p; return Fail(QLB200_ERR_UNSUPPORTED, err); }
  if (ctx != nullptr) {   // ctx == NULL: host-only plan (stats / partition queries, no device tables)
    int rc = FinishPlan(ctx, p);
    if (rc != QLB200_OK) { p->d.Free(); delete p; return rc; }
  }
  *out = p;
  return QLB200_OK;
}

int qlb200_plan_create_raw(qlb200_ctx *ctx, int dtype, uint32_t flags, int32_t a_rank, const int32_t *a_perm,
                           uint64_t na, const uint32_t *a_shape, const uint64_t *a_off, int32_t b_rank,
                           const int32_t *b_perm, uint64_t nb, const uint32_t *b_shape, const uint64_t *b_off,
                           uint64_t ntask, const qlb200_task *tasks, uint64_t c_elems, qlb200_plan **out) {
  if (!out || !a_shape || !b_shape || !a_off || !b_off || (ntask && !tasks)) return Fail(QLB200_ERR_ARG, "null argument");
  if (dtype != QLB200_F64 && dtype != QLB200_C64) return Fail(QLB200_ERR_ARG, "bad dtype");
  if (a_rank < 1 || a_rank > QLB200_MAX_RANK || b_rank < 1 || b_rank > QLB200_MAX_RANK) return Fail(QLB200_ERR_ARG, "bad rank");
  auto total = [](int rank, uint64_t n, const uint32_t *shape, const uint64_t *off) {
    uint64_t e = 0;
    for (uint64_t b = 0; b < n; ++b) {
      uint64_t sz = 1;
      for (int i = 0; i < rank; ++i) sz *= shape[b * rank + i];
      e = std::max(e, off[b] + sz);
    }
    return e;
  };
  std::vector<qlb200_task> st(tasks, tasks + ntask);
  std::stable_sort(st.begin(), st.end(), [](const qlb200_task &x, const qlb200_task &y) {
    if (x.c_off != y.c_off) return x.c_off < y.c_off;
    return x.first > y.first;
  });
  for (auto &t : st) t.c_ord = 0;   // groups are delimited by c_off here
  {
    uint32_t ord = 0;
    for (size_t i = 0; i < st.size(); ++i) { if (i > 0 && st[i].c_off != st[i - 1].c_off) ++ord; st[i].c_ord = ord; }
  }
  qlb200_plan *p = new (std::nothrow) qlb200_plan();
  if (!p) return Fail(QLB200_ERR_NOMEM, "out of memory");
  p->ctx = ctx;
  SetPlanKnobs(ctx, &flags, &p->h);
  std::string err = BuildPlanHost(dtype, flags, 0, a_rank, a_perm, na, a_shape, a_off, total(a_rank, na, a_shape, a_off),
                                  b_rank, b_perm, nb, b_shape, b_off, total(b_rank, nb, b_shape, b_off), st, c_elems, &p->h);
  if (!err.empty()) { delete p; return Fail(QLB200_ERR_UNSUPPORTED, err); }
  if (ctx != nullptr) {   // ctx == NULL: host-only plan (stats / partition queries, no device tables)
    int rc = FinishPlan(ctx, p);
    if (rc != QLB200_OK) { p->d.Free(); delete p; return rc; }
  }
  *out = p;
  return QLB200_OK;
}

void qlb200_plan_destroy(qlb200_plan *p) {
  if (!p) return;
  if (!p->ctx) { delete p; return; }
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  p->d.Free();
  if (p->d_acc_ranges) cudaFree(p->d_acc_ranges);
  delete p;
}

int qlb200_plan_partition(qlb200_plan *p, int32_t world, int32_t rank) {
  if (!p || world < 1 || rank < 0 || rank >= world) return Fail(QLB200_ERR_ARG, "bad world/rank");
  PartitionRows(&p->h, world, rank);
  std::string err = BuildTiles(&p->h);
  if (!err.empty()) return Fail(QLB200_ERR_UNSUPPORTED, err);
  if (!p->ctx) return QLB200_OK;
  QL_CUDA(cudaSetDevice(p->ctx->device));
  if (!p->h.perm_blks_all.empty()) {       // the permute tables shrink with the share
    QL_CUDA(cudaStreamSynchronize(p->ctx->stream));
    cudaFree(p->d.perm_blks); cudaFree(p->d.perm_tile_base);
    p->d.perm_blks = nullptr; p->d.perm_tile_base = nullptr;
    int rc = Upload(p->h.perm_blks, &p->d.perm_blks, p->ctx->stream);
    if (rc == QLB200_OK) rc = Upload(p->h.perm_tile_base, &p->d.perm_tile_base, p->ctx->stream);
    if (rc != QLB200_OK) return rc;
  }
  return UploadGemmTables(p);
}
uint64_t qlb200_plan_c_range_count(const qlb200_plan *p) {
  uint64_t n = 0;
  for (const auto &g : p->h.part_groups) if (g.row_end > g.row_begin) ++n;
  return n;
}
int qlb200_plan_c_ranges(const qlb200_plan *p, uint64_t *off, uint64_t *len) {
  uint64_t i = 0;
  for (const auto &g : p->h.part_groups) {
    if (g.row_end <= g.row_begin) continue;
    off[i] = g.c_off + uint64_t(g.row_begin) * g.n;
    len[i] = uint64_t(g.row_end - g.row_begin) * g.n;
    ++i;
  }
  return QLB200_OK;
}
int qlb200_plan_get_stats(const qlb200_plan *p, qlb200_plan_stats *out) {
  if (!p || !out) return Fail(QLB200_ERR_ARG, "null argument");
  std::memset(out, 0, sizeof(*out));
  out->flops = p->h.flops;
  out->ntask = p->h.tasks.size();
  out->ngroup = p->h.groups.size();
  out->ntile_dmma = p->h.tiles.size();
  out->nrow_skinny = p->h.items.size();
  out->permute_elems_a = p->h.permute_elems_a;
  out->permute_elems_b = p->h.permute_elems_b;
  out->workspace_bytes = WsBytes(p);
  out->gemm_read_bytes = p->h.gemm_read_bytes;
  out->gemm_write_bytes = p->h.gemm_write_bytes;
  return QLB200_OK;
}

int qlb200_plan_operand_block(const qlb200_plan *p, int which, uint64_t ord, uint64_t *ws_off) {
  if (!p || (which != 0 && which != 1)) return Fail(QLB200_ERR_ARG, "bad argument");
  const std::vector<uint64_t> &v = which == 0 ? p->h.ws_off_a : p->h.ws_off_b;
  if (ord >= v.size()) return Fail(QLB200_ERR_ARG, "block ordinal out of range");
  if (v[ord] == ~0ull) return 0;
  if (ws_off) *ws_off = v[ord];
  return 1;
}

int qlb200_plan_read_workspace(qlb200_ctx *ctx, qlb200_plan *p, int which, uint64_t elem_off, uint64_t elems, void *dst_host) {
  if (!ctx || !p || !dst_host || (which != 0 && which != 1)) return Fail(QLB200_ERR_ARG, "bad argument");
  if (p->ctx != ctx) return Fail(QLB200_ERR_ARG, "plan belongs to another context");
  const size_t es = ElemSize(p->h.dtype);
  const uint64_t have = which == 0 ? p->h.ws_a_elems : p->h.ws_b_elems;
  if (elem_off > have || elems > have - elem_off) return Fail(QLB200_ERR_ARG, "range outside the permuted workspace");
  if (WsBytes(p) > ctx->ws_bytes) return Fail(QLB200_ERR_ARG, "workspace not populated: call qlb200_execute_permute first");
  QL_CUDA(cudaSetDevice(ctx->device));
  const char *base = static_cast<const char *>(ctx->ws) + (which == 0 ? 0 : Align256(p->h.ws_a_elems * es));
  QL_CUDA(cudaMemcpyAsync(dst_host, base + elem_off * es, elems * es, cudaMemcpyDeviceToHost, ctx->stream));
  QL_CUDA(cudaStreamSynchronize(ctx->stream));
  return QLB200_OK;
}

uint64_t qlb200_plan_units(const qlb200_plan *p, uint64_t cap, qlb200_unit *out, uint32_t *tile_rows, uint32_t *tile_cols,
                           uint32_t *stage_k) {
  if (!p) return 0;
  const bool four_m = (p->h.flags & QLB200_PLAN_CPLX_4M) != 0;
  uint32_t BM, BN, BK;
  if (p->h.dtype == QLB200_C64) { BM = kWsBM; BN = four_m ? kWsBN : kWs3mBN; BK = four_m ? kWsBK : kWs3mBK; }
  else { BM = kWsRealBM; BN = kWsRealBN; BK = kWsRealBK; }
  if (tile_rows) *tile_rows = BM;
  if (tile_cols) *tile_cols = BN;
  if (stage_k) *stage_k = BK;
  const uint64_t n = p->h.tiles.size();
  for (uint64_t i = 0; i < n && i < cap && out; ++i) {
    const GemmTile &t = p->h.tiles[i];
    const GemmGroup &g = p->h.part_groups[t.group];
    qlb200_unit &u = out[i];
    u.group = t.group; u.tm = t.tm; u.tn = t.tn; u.s_begin = t.s_begin; u.s_end = t.s_end; u.split = t.split; u.nsplit = t.nsplit;
    u.rows = std::min<uint32_t>(BM, g.row_end - g.row_begin - uint32_t(t.tm) * BM);
    u.cols = std::min<uint32_t>(BN, g.n - uint32_t(t.tn) * BN);
  }
  return n;
}

int qlb200_shard_sector_flops(const qlb200_match *m, int32_t axis, int dtype, double *cost) {
  if (!m || !cost || axis < 0 || axis >= m->m.a.rank) return Fail(QLB200_ERR_ARG, "bad argument");
  const Shell &a = m->m.a;
  const double f = dtype == QLB200_C64 ? 8.0 : 2.0;
  for (const qlb200_task &t : m->m.tasks)
    cost[a.coors[uint64_t(t.a_ord) * a.rank + axis]] += f * double(t.m) * double(t.k) * double(t.n);
  return QLB200_OK;
}

uint64_t qlb200_plan_items(const qlb200_plan *p, uint64_t cap, qlb200_item *out) {
  if (!p) return 0;
  const uint64_t n = p->h.items.size();
  for (uint64_t i = 0; i < n && i < cap && out; ++i) {
    const SkinnyItem &it = p->h.items[i];
    const GemmGroup &g = p->h.part_groups[it.group];
    const uint32_t sub_rows = uint32_t(kSkinnyElems) / g.n;       // as the kernel computes it
    out[i].group = it.group; out[i].row0 = it.row0; out[i].n = g.n;
    out[i].rows = std::min<uint32_t>(sub_rows * p->h.skinny_sub, g.row_end - it.row0);
  }
  return n;
}

uint64_t qlb200_plan_segments(const qlb200_plan *p, uint64_t cap, uint32_t *seg_out) {
  if (!p) return 0;
  for (uint64_t i = 0; i < p->h.seg.size() && i < cap && seg_out; ++i) seg_out[i] = p->h.seg[i];
  return p->h.seg.size();
}

// ---- execution ---------------------------------------------------------------------------------
// Checks shared by every execute entry point: a host-only plan has no device tables, a plan of another context
// lives on another device / stream.
static int CheckExec(const qlb200_ctx *ctx, const qlb200_plan *p, const void *A, const void *B) {
  if (!ctx || !p || !A || !B) return Fail(QLB200_ERR_ARG, "null argument");
  if (p->ctx == nullptr) return Fail(QLB200_ERR_ARG, "host-only plan (created without a context) cannot execute");
  if (p->ctx != ctx) return Fail(QLB200_ERR_ARG, "plan belongs to another context");
  return QLB200_OK;
}

static int ResolveWorkspace(qlb200_ctx *ctx, qlb200_plan *p, void **wsA, void **wsB, void **partials = nullptr) {
  const size_t es = ElemSize(p->h.dtype);
  int rc = EnsureArena(ctx, &ctx->ws, &ctx->ws_bytes, WsBytes(p));
  if (rc != QLB200_OK) return rc;
  *wsA = ctx->ws;
  *wsB = static_cast<char *>(ctx->ws) + Align256(p->h.ws_a_elems * es);
  if (partials) *partials = static_cast<char *>(*wsB) + Align256(p->h.ws_b_elems * es) + p->partials_shift;
  return QLB200_OK;
}

int qlb200_execute_permute(qlb200_ctx *ctx, qlb200_plan *p, const void *A, const void *B) {
  int ok = CheckExec(ctx, p, A, B);
  if (ok != QLB200_OK) return ok;
  QL_CUDA(cudaSetDevice(ctx->device));
  void *wa, *wb;
  int rc = ResolveWorkspace(ctx, p, &wa, &wb);
  if (rc != QLB200_OK) return rc;
  ctx->launches = 0;
  const uint32_t ntiles = p->h.perm_tile_base.back();
  if (ntiles > 0) {
    QL_CUDA(LaunchPermute(p->h.dtype, p->d.perm_blks, p->d.perm_tile_base, static_cast<uint32_t>(p->h.perm_blks.size()),
                          ntiles, A, B, wa, wb, ctx->num_sms, ctx->stream));
    ctx->launches += 1; ctx->total_launches += 1;
  }
  return QLB200_OK;
}

static int ExecuteGemm(qlb200_ctx *ctx, qlb200_plan *p, const void *A, const void *B, void *const *c_out, uint32_t n_out,
                       uint32_t mcast = 0, const void *c_in = nullptr) {
  int ok = CheckExec(ctx, p, A, B);
  if (ok != QLB200_OK) return ok;
  if (!c_out || n_out < 1 || n_out > uint32_t(kMaxOut)) return Fail(QLB200_ERR_ARG, "bad output list");
  for (uint32_t d = 0; d < n_out; ++d) if (!c_out[d]) return Fail(QLB200_ERR_ARG, "null output pointer");
  QL_CUDA(cudaSetDevice(ctx->device));
  void *wa, *wb, *parts;
  int rc = ResolveWorkspace(ctx, p, &wa, &wb, &parts);
  if (rc != QLB200_OK) return rc;
  ctx->launches = 0;
  GemmParams gp = MakeParams(p, A, B, wa, wb, parts, c_out, n_out, mcast);
  if (p->accum) {
    if (n_out != 1 || mcast) return Fail(QLB200_ERR_UNSUPPORTED, "an accumulate plan writes one local output (use qlb200_execute_accum)");
    gp.accum = 1; gp.c_in = c_in;
    gp.alpha_re = p->alpha[0]; gp.alpha_im = p->alpha[1]; gp.beta_re = p->beta[0]; gp.beta_im = p->beta[1];
  }
  // A plan that has both DMMA tiles and a handful of narrow-pair items (the small sectors of a GEMM-shaped step):
  // the narrow kernel is a 10-15 us latency-bound launch on a few SMs.  It is forked onto the side stream FIRST, so the
  // persistent DMMA CTAs fill the remaining SMs at once and the narrow kernel costs no time of its own (two in-order
  // launches on one stream would serialise).  The two kernels write disjoint output blocks.  Fork / join by events is
  // legal under stream capture (qlb200_graph_*).
  const bool beside = gp.ntiles > 0 && gp.nitems > 0 && gp.nitems <= uint32_t(ctx->num_sms) && ctx->side != nullptr;
  if (beside) {
    QL_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    QL_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    QL_CUDA(LaunchGemmSkinny(p->h.dtype, gp, ctx->num_sms, ctx->side));
    QL_CUDA(cudaEventRecord(ctx->ev_join, ctx->side));
    ctx->launches += 1; ctx->total_launches += 1;
  }
  if (gp.ntiles > 0) {
    if (p->h.dtype == QLB200_C64) {
      QL_CUDA(LaunchGemmWsCplx(gp, !(p->h.flags & QLB200_PLAN_CPLX_4M), ctx->num_sms, ctx->stream));
    } else {
      QL_CUDA(LaunchGemmWsReal(gp, ctx->num_sms, ctx->stream));
    }
    ctx->launches += 1; ctx->total_launches += 1;
  }
  if (beside) {
    QL_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  } else if (gp.nitems > 0) {
    QL_CUDA(LaunchGemmSkinny(p->h.dtype, gp, ctx->num_sms, ctx->stream));
    ctx->launches += 1; ctx->total_launches += 1;
  }
  return QLB200_OK;
}

int qlb200_execute_gemm(qlb200_ctx *ctx, qlb200_plan *p, const void *A, const void *B, void *C) {
  void *outs[1] = {C};
  return ExecuteGemm(ctx, p, A, B, outs, 1);
}

int qlb200_execute_bcast(qlb200_ctx *ctx, qlb200_plan *p, const void *A, const void *B, void *const *C_peers, int32_t npeers) {
  if (!ctx || !p || !A || !B || !C_peers) return Fail(QLB200_ERR_ARG, "null argument");
  if (npeers < 1 || npeers > kMaxOut) return Fail(QLB200_ERR_ARG, "npeers must be 1..8");
  for (int32_t d = 0; d < npeers; ++d) if (!C_peers[d]) return Fail(QLB200_ERR_ARG, "null output pointer");
  if (p->ctx != ctx) return Fail(QLB200_ERR_ARG, "plan belongs to another context");
  int rc = qlb200_execute_permute(ctx, p, A, B);
  if (rc != QLB200_OK) return rc;
  const uint64_t l0 = ctx->launches;
  rc = ExecuteGemm(ctx, p, A, B, C_peers, static_cast<uint32_t>(npeers));
  if (rc != QLB200_OK) return rc;
  ctx->launches += l0;
  return QLB200_OK;
}

int qlb200_execute_mcast(qlb200_ctx *ctx, qlb200_plan *p, const void *A, const void *B, void *C_multicast) {
  if (!ctx || !p || !A || !B || !C_multicast) return Fail(QLB200_ERR_ARG, "null argument");
  if (p->ctx != ctx) return Fail(QLB200_ERR_ARG, "plan belongs to another context");
  int rc = qlb200_execute_permute(ctx, p, A, B);
  if (rc != QLB200_OK) return rc;
  const uint64_t l0 = ctx->launches;
  void *out[1] = {C_multicast};
  rc = ExecuteGemm(ctx, p, A, B, out, 1, 1);
  if (rc != QLB200_OK) return rc;
  ctx->launches += l0;
  return QLB200_OK;
}

int qlb200_fanout_copy(qlb200_ctx *ctx, const void *src, uint64_t byte_off, uint64_t bytes, void *const *dst_peers, int32_t npeers,
                       void *dst_multicast) {
  if (!ctx || !src) return Fail(QLB200_ERR_ARG, "null argument");
  if ((dst_multicast == nullptr) == (dst_peers == nullptr || npeers < 1)) return Fail(QLB200_ERR_ARG, "give either peer pointers or a multicast pointer");
  if (!dst_multicast && npeers > kMaxOut) return Fail(QLB200_ERR_ARG, "npeers must be 1..8");
  if ((byte_off | bytes) & 15u) return Fail(QLB200_ERR_ARG, "offset and length must be multiples of 16 bytes");
  QL_CUDA(cudaSetDevice(ctx->device));
  void *dst[kMaxOut] = {};
  uint32_t nd = 1;
  if (dst_multicast) dst[0] = static_cast<char *>(dst_multicast) + byte_off;
  else { nd = uint32_t(npeers); for (uint32_t i = 0; i < nd; ++i) { if (!dst_peers[i]) return Fail(QLB200_ERR_ARG, "null peer pointer"); dst[i] = static_cast<char *>(dst_peers[i]) + byte_off; } }
  QL_CUDA(LaunchFanOutCopy(static_cast<const char *>(src) + byte_off, bytes, dst, nd, dst_multicast != nullptr, ctx->num_sms, ctx->stream));
  ctx->launches = 1; ctx->total_launches += 1;
  return QLB200_OK;
}

int qlb200_plan_remap_output(qlb200_plan *p, uint64_t n, const uint64_t *from_off, const uint64_t *to_off) {
  if (!p || (n && (!from_off || !to_off))) return Fail(QLB200_ERR_ARG, "null argument");
  std::vector<std::pair<uint64_t, uint64_t>> map(n);
  for (uint64_t i = 0; i < n; ++i) map[i] = {from_off[i], to_off[i]};
  std::sort(map.begin(), map.end());
  for (size_t gi = 0; gi < p->h.groups.size(); ++gi) {
    const uint64_t off = p->h.groups[gi].c_off;
    auto it = std::lower_bound(map.begin(), map.end(), std::make_pair(off, uint64_t(0)));
    if (it == map.end() || it->first != off) return Fail(QLB200_ERR_ARG, "output block offset missing from the remap table");
    p->h.groups[gi].c_off = it->second;
    p->h.part_groups[gi].c_off = it->second;
  }
  if (!p->ctx) return QLB200_OK;
  QL_CUDA(cudaSetDevice(p->ctx->device));
  return UploadGemmTables(p);
}

struct qlb200_graph {
  qlb200_ctx *ctx = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
};

int qlb200_graph_begin(qlb200_ctx *ctx) {
  if (!ctx) return Fail(QLB200_ERR_ARG, "null argument");
  QL_CUDA(cudaSetDevice(ctx->device));
  QL_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
  return QLB200_OK;
}
int qlb200_graph_end(qlb200_ctx *ctx, qlb200_graph **out) {
  if (!ctx || !out) return Fail(QLB200_ERR_ARG, "null argument");
  cudaGraph_t graph = nullptr;
  QL_CUDA(cudaStreamEndCapture(ctx->stream, &graph));
  if (!graph) return Fail(QLB200_ERR_CUDA, "stream capture produced no graph");
  cudaGraphExec_t exec = nullptr;
  cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
  if (e != cudaSuccess) { cudaGraphDestroy(graph); return Fail(QLB200_ERR_CUDA, CudaErr("cudaGraphInstantiate", e)); }
  qlb200_graph *g = new qlb200_graph;
  g->ctx = ctx; g->graph = graph; g->exec = exec;
  ++ctx->graphs_alive;
  *out = g;
  return QLB200_OK;
}
int qlb200_graph_launch(qlb200_ctx *ctx, qlb200_graph *g) {
  if (!ctx || !g || !g->exec) return Fail(QLB200_ERR_ARG, "null argument");
  if (g->ctx != ctx) return Fail(QLB200_ERR_ARG, "graph was captured on another context");
  QL_CUDA(cudaGraphLaunch(g->exec, ctx->stream));
  return QLB200_OK;
}
void qlb200_graph_destroy(qlb200_graph *g) {
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  if (g->ctx && --g->ctx->graphs_alive == 0) {      // nothing can replay into the outgrown arenas any more
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);
    for (void *r : g->ctx->retired) cudaFree(r);
    g->ctx->retired.clear();
  }
  delete g;
}

int qlb200_ipc_export(qlb200_ctx *ctx, const void *dev_ptr, unsigned char *handle64) {
  if (!ctx || !dev_ptr || !handle64) return Fail(QLB200_ERR_ARG, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  QL_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  QL_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
  std::memcpy(handle64, &h, 64);
  return QLB200_OK;
}
int qlb200_ipc_open(qlb200_ctx *ctx, const unsigned char *handle64, void **peer_ptr) {
  if (!ctx || !handle64 || !peer_ptr) return Fail(QLB200_ERR_ARG, "null argument");
  QL_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  QL_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return QLB200_OK;
}
int qlb200_ipc_close(qlb200_ctx *ctx, void *peer_ptr) {
  if (!ctx || !peer_ptr) return Fail(QLB200_ERR_ARG, "null argument");
  QL_CUDA(cudaSetDevice(ctx->device));
  QL_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return QLB200_OK;
}

int qlb200_execute(qlb200_ctx *ctx, qlb200_plan *p, const void *A, const void *B, void *C, int mem_kind) {
  if (!ctx || !p || !A || !B || !C) return Fail(QLB200_ERR_ARG, "null argument");
  if (p->ctx == nullptr) return Fail(QLB200_ERR_ARG, "host-only plan (created without a context) cannot execute");
  if (p->ctx != ctx) return Fail(QLB200_ERR_ARG, "plan belongs to another context");
  QL_CUDA(cudaSetDevice(ctx->device));
  const size_t es = ElemSize(p->h.dtype);
  const void *dA = A, *dB = B;
  void *dC = C;
  if (mem_kind == QLB200_MEM_HOST) {
    const size_t ab = Align256(p->h.a_elems * es), bb = Align256(p->h.b_elems * es), cb = Align256(p->h.c_elems * es);
    int rc = EnsureArena(ctx, &ctx->stage, &ctx->stage_bytes, ab + bb + cb);
    if (rc != QLB200_OK) return rc;
    char *base = static_cast<char *>(ctx->stage);
    QL_CUDA(cudaMemcpyAsync(base, A, p->h.a_elems * es, cudaMemcpyHostToDevice, ctx->stream));
    QL_CUDA(cudaMemcpyAsync(base + ab, B, p->h.b_elems * es, cudaMemcpyHostToDevice, ctx->stream));
    dA = base; dB = base + ab; dC = base + ab + bb;
  } else if (mem_kind != QLB200_MEM_DEVICE) {
    return Fail(QLB200_ERR_ARG, "bad mem_kind");
  }
  int rc = qlb200_execute_permute(ctx, p, dA, dB);
  if (rc != QLB200_OK) return rc;
  const uint64_t l0 = ctx->launches;
  rc = qlb200_execute_gemm(ctx, p, dA, dB, dC);
  if (rc != QLB200_OK) return rc;
  ctx->launches += l0;
  if (mem_kind == QLB200_MEM_HOST) {
    // only the ranges this plan writes are copied back (whole C unless partitioned)
    bool whole = true;
    for (size_t i = 0; i < p->h.groups.size(); ++i)
      if (p->h.part_groups[i].row_begin != 0 || p->h.part_groups[i].row_end != p->h.groups[i].m) { whole = false; break; }
    if (whole) {
      QL_CUDA(cudaMemcpyAsync(C, dC, p->h.c_elems * es, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      for (const auto &g : p->h.part_groups) {
        if (g.row_end <= g.row_begin) continue;
        const size_t o = (g.c_off + size_t(g.row_begin) * g.n) * es, l = size_t(g.row_end - g.row_begin) * g.n * es;
        QL_CUDA(cudaMemcpyAsync(static_cast<char *>(C) + o, static_cast<char *>(dC) + o, l, cudaMemcpyDeviceToHost, ctx->stream));
      }
    }
    QL_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return QLB200_OK;
}

// ---- plan splitting and the host pipeline -----------------------------------------------------------
int qlb200_plan_split(const qlb200_plan *p, int by, int32_t nparts, const double *cum_frac, qlb200_plan **parts_out,
                      uint64_t *bounds_out) {
  if (!p || !cum_frac || !parts_out || !bounds_out || nparts < 1) return Fail(QLB200_ERR_ARG, "bad argument");
  if (by != QLB200_SPLIT_BY_A && by != QLB200_SPLIT_BY_B && by != QLB200_SPLIT_BY_C) return Fail(QLB200_ERR_ARG, "bad split kind");
  if (p->accum) return Fail(QLB200_ERR_UNSUPPORTED, "accumulate plans are not split");
  if (!p->h.perm_blks.empty()) return Fail(QLB200_ERR_UNSUPPORTED, "plan has blocks that go through the permute kernel");
  const PlanHost &h = p->h;
  const size_t ng = h.groups.size();
  // key of a group: the end of the last operand element it needs (arrival splits) or of its output block (output split)
  std::vector<uint64_t> key(ng, 0);
  const uint64_t total = by == QLB200_SPLIT_BY_A ? h.a_elems : by == QLB200_SPLIT_BY_B ? h.b_elems : h.c_elems;
  for (size_t g = 0; g < ng; ++g) {
    const GemmGroup &gg = h.groups[g];
    if (by == QLB200_SPLIT_BY_C) { key[g] = gg.c_off + uint64_t(gg.m) * gg.n; continue; }
    for (uint32_t t = gg.task_begin; t < gg.task_end; ++t) {
      const GemmTask &tk = h.tasks[t];
      const uint64_t end = by == QLB200_SPLIT_BY_A ? tk.a_off + uint64_t(gg.m) * tk.k : tk.b_off + uint64_t(tk.k) * gg.n;
      key[g] = std::max(key[g], end);
    }
  }
  // cut positions: the smallest group key at or above the wanted cumulative share (a key is a block end, so a cut never
  // separates a block from itself)
  std::vector<uint64_t> sorted_keys(key);
  std::sort(sorted_keys.begin(), sorted_keys.end());
  bounds_out[0] = 0;
  for (int32_t i = 0; i < nparts; ++i) {
    uint64_t want = i + 1 == nparts ? total : uint64_t(cum_frac[i] * double(total));
    auto it = std::lower_bound(sorted_keys.begin(), sorted_keys.end(), want);
    uint64_t cut = i + 1 == nparts ? total : (it == sorted_keys.end() ? total : *it);
    bounds_out[i + 1] = std::max(cut, bounds_out[i]);
  }
  for (int32_t i = 0; i < nparts; ++i) parts_out[i] = nullptr;
  for (int32_t i = 0; i < nparts; ++i) {
    qlb200_plan *q = new (std::nothrow) qlb200_plan();
    if (!q) return Fail(QLB200_ERR_NOMEM, "out of memory");
    q->ctx = p->ctx;
    q->h = p->h;
    for (size_t g = 0; g < ng; ++g) {
      const bool mine = key[g] > bounds_out[i] && key[g] <= bounds_out[i + 1];
      if (!mine) q->h.part_groups[g].row_end = q->h.part_groups[g].row_begin;     // empty row range: no tiles, no items
    }
    std::string err = BuildTiles(&q->h);
    int rc = err.empty() ? QLB200_OK : Fail(QLB200_ERR_UNSUPPORTED, err);
    if (rc == QLB200_OK && q->ctx != nullptr) {
      cudaSetDevice(q->ctx->device);
      rc = UploadGemmTables(q);
    }
    if (rc != QLB200_OK) {
      q->d.Free(); delete q;
      for (int32_t j = 0; j < i; ++j) { qlb200_plan_destroy(parts_out[j]); parts_out[j] = nullptr; }
      return rc;
    }
    parts_out[i] = q;
  }
  return QLB200_OK;
}

struct qlb200_hostpipe {
  qlb200_ctx *ctx = nullptr;
  int first_streams = QLB200_SPLIT_BY_B;
  int dtype = 0;
  std::vector<qlb200_plan *> in_parts, out_parts;
  std::vector<uint64_t> in_bounds, out_bounds;
  std::vector<cudaEvent_t> ev_in, ev_out;
  cudaEvent_t ev_start = nullptr, ev_done = nullptr, ev_alt_in = nullptr, ev_alt_out = nullptr, ev_mid = nullptr;
  uint64_t launches = 0;
};

// Runs one part on `on` (the context's main stream or its alternate stream): the kernels only see ctx->stream.
static int ExecutePartOn(qlb200_ctx *ctx, cudaStream_t on, qlb200_plan *q, const void *A, const void *B, void *C) {
  cudaStream_t saved = ctx->stream;
  ctx->stream = on;
  const int rc = qlb200_execute_gemm(ctx, q, A, B, C);
  ctx->stream = saved;
  return rc;
}

void qlb200_hostpipe_destroy(qlb200_hostpipe *hp) {
  if (!hp) return;
  if (hp->ctx) { cudaSetDevice(hp->ctx->device); cudaStreamSynchronize(hp->ctx->stream); cudaStreamSynchronize(hp->ctx->copy); }
  for (qlb200_plan *q : hp->in_parts) qlb200_plan_destroy(q);
  for (qlb200_plan *q : hp->out_parts) qlb200_plan_destroy(q);
  for (cudaEvent_t e : hp->ev_in) cudaEventDestroy(e);
  for (cudaEvent_t e : hp->ev_out) cudaEventDestroy(e);
  for (cudaEvent_t e : {hp->ev_start, hp->ev_done, hp->ev_alt_in, hp->ev_alt_out, hp->ev_mid}) if (e) cudaEventDestroy(e);
  delete hp;
}

int qlb200_hostpipe_create(qlb200_ctx *ctx, const qlb200_plan *first, int first_streams, int32_t nparts_in, const double *cum_in,
                           const qlb200_plan *last, int32_t nparts_out, const double *cum_out, qlb200_hostpipe **out) {
  if (!ctx || !first || !last || !out || !cum_in || !cum_out || nparts_in < 1 || nparts_out < 1) return Fail(QLB200_ERR_ARG, "bad argument");
  if (first->ctx != ctx || last->ctx != ctx) return Fail(QLB200_ERR_ARG, "plan belongs to another context");
  if (first_streams != QLB200_SPLIT_BY_A && first_streams != QLB200_SPLIT_BY_B) return Fail(QLB200_ERR_ARG, "the first step streams operand A or B");
  if (first->h.dtype != last->h.dtype) return Fail(QLB200_ERR_ARG, "plans of different element types");
  QL_CUDA(cudaSetDevice(ctx->device));
  qlb200_hostpipe *hp = new (std::nothrow) qlb200_hostpipe();
  if (!hp) return Fail(QLB200_ERR_NOMEM, "out of memory");
  hp->ctx = ctx; hp->first_streams = first_streams; hp->dtype = first->h.dtype;
  hp->in_parts.assign(nparts_in, nullptr); hp->in_bounds.assign(nparts_in + 1, 0);
  hp->out_parts.assign(nparts_out, nullptr); hp->out_bounds.assign(nparts_out + 1, 0);
  int rc = qlb200_plan_split(first, first_streams, nparts_in, cum_in, hp->in_parts.data(), hp->in_bounds.data());
  if (rc == QLB200_OK) rc = qlb200_plan_split(last, QLB200_SPLIT_BY_C, nparts_out, cum_out, hp->out_parts.data(), hp->out_bounds.data());
  if (rc != QLB200_OK) { qlb200_hostpipe_destroy(hp); return rc; }
  // concurrent parts need disjoint split-K partial-tile regions, and the arena must have its final size before any part
  // runs on the alternate stream (growth synchronises only the main stream)
  {
    uint64_t shift = 0, need = 0;
    const size_t es = ElemSize(hp->dtype);
    for (auto *vec : {&hp->in_parts, &hp->out_parts}) {
      shift = 0;
      for (qlb200_plan *q : *vec) {
        q->partials_shift = shift;
        shift += Align256(q->h.n_part_slots * q->h.part_slot_elems * es);
        need = std::max<uint64_t>(need, WsBytes(q));
      }
    }
    rc = EnsureArena(ctx, &ctx->ws, &ctx->ws_bytes, need);
    if (rc != QLB200_OK) { qlb200_hostpipe_destroy(hp); return rc; }
  }
  auto mk = [](cudaEvent_t *e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming); };
  cudaError_t e = mk(&hp->ev_start);
  if (e == cudaSuccess) e = mk(&hp->ev_done);
  if (e == cudaSuccess) e = mk(&hp->ev_alt_in);
  if (e == cudaSuccess) e = mk(&hp->ev_alt_out);
  if (e == cudaSuccess) e = mk(&hp->ev_mid);
  hp->ev_in.assign(nparts_in, nullptr); hp->ev_out.assign(nparts_out, nullptr);
  for (auto &ev : hp->ev_in) if (e == cudaSuccess) e = mk(&ev);
  for (auto &ev : hp->ev_out) if (e == cudaSuccess) e = mk(&ev);
  if (e != cudaSuccess) { qlb200_hostpipe_destroy(hp); return Fail(QLB200_ERR_CUDA, CudaErr("cudaEventCreate", e)); }
  *out = hp;
  return QLB200_OK;
}

int qlb200_hostpipe_begin(qlb200_ctx *ctx, qlb200_hostpipe *hp, const void *in_host, void *in_dev, const void *other_dev, void *c_dev) {
  if (!ctx || !hp || !in_host || !in_dev || !other_dev || !c_dev) return Fail(QLB200_ERR_ARG, "null argument");
  if (hp->ctx != ctx) return Fail(QLB200_ERR_ARG, "pipe belongs to another context");
  QL_CUDA(cudaSetDevice(ctx->device));
  const size_t es = ElemSize(hp->dtype);
  hp->launches = 0;
  // the copy stream starts where the compute stream is now (the previous apply may still read in_dev)
  QL_CUDA(cudaEventRecord(hp->ev_start, ctx->stream));
  QL_CUDA(cudaStreamWaitEvent(ctx->copy, hp->ev_start, 0));
  const size_t np = hp->in_parts.size();
  for (size_t i = 0; i < np; ++i) {
    const uint64_t lo = hp->in_bounds[i], hi = hp->in_bounds[i + 1];
    if (hi > lo)
      QL_CUDA(cudaMemcpyAsync(static_cast<char *>(in_dev) + lo * es, static_cast<const char *>(in_host) + lo * es, (hi - lo) * es,
                              cudaMemcpyHostToDevice, ctx->copy));
    QL_CUDA(cudaEventRecord(hp->ev_in[i], ctx->copy));
  }
  // parts alternate between the main and the alternate stream: part i + 1 fills the SMs part i's last units leave idle
  QL_CUDA(cudaStreamWaitEvent(ctx->alt, hp->ev_start, 0));
  bool used_alt = false;
  for (size_t i = 0; i < np; ++i) {
    cudaStream_t on = (i & 1) ? ctx->alt : ctx->stream;
    QL_CUDA(cudaStreamWaitEvent(on, hp->ev_in[i], 0));
    qlb200_plan *q = hp->in_parts[i];
    if (q->h.tiles.empty() && q->h.items.empty()) continue;
    const void *A = hp->first_streams == QLB200_SPLIT_BY_A ? in_dev : other_dev;
    const void *B = hp->first_streams == QLB200_SPLIT_BY_A ? other_dev : in_dev;
    int rc = ExecutePartOn(ctx, on, q, A, B, c_dev);
    if (rc != QLB200_OK) return rc;
    hp->launches += ctx->launches;
    used_alt = used_alt || (i & 1);
  }
  if (used_alt) {      // whatever follows on the main stream needs every part
    QL_CUDA(cudaEventRecord(hp->ev_alt_in, ctx->alt));
    QL_CUDA(cudaStreamWaitEvent(ctx->stream, hp->ev_alt_in, 0));
  }
  return QLB200_OK;
}

int qlb200_hostpipe_end(qlb200_ctx *ctx, qlb200_hostpipe *hp, const void *a_dev, const void *b_dev, void *c_dev, void *out_host) {
  if (!ctx || !hp || !a_dev || !b_dev || !c_dev || !out_host) return Fail(QLB200_ERR_ARG, "null argument");
  if (hp->ctx != ctx) return Fail(QLB200_ERR_ARG, "pipe belongs to another context");
  QL_CUDA(cudaSetDevice(ctx->device));
  const size_t es = ElemSize(hp->dtype);
  const size_t np = hp->out_parts.size();
  QL_CUDA(cudaEventRecord(hp->ev_mid, ctx->stream));          // the last step's operands are ready here
  QL_CUDA(cudaStreamWaitEvent(ctx->alt, hp->ev_mid, 0));
  for (size_t i = 0; i < np; ++i) {
    cudaStream_t on = (i & 1) ? ctx->alt : ctx->stream;
    qlb200_plan *q = hp->out_parts[i];
    if (!(q->h.tiles.empty() && q->h.items.empty())) {
      int rc = ExecutePartOn(ctx, on, q, a_dev, b_dev, c_dev);
      if (rc != QLB200_OK) return rc;
      hp->launches += ctx->launches;
    }
    QL_CUDA(cudaEventRecord(hp->ev_out[i], on));
    QL_CUDA(cudaStreamWaitEvent(ctx->copy, hp->ev_out[i], 0));
    const uint64_t lo = hp->out_bounds[i], hi = hp->out_bounds[i + 1];
    if (hi > lo)
      QL_CUDA(cudaMemcpyAsync(static_cast<char *>(out_host) + lo * es, static_cast<const char *>(c_dev) + lo * es, (hi - lo) * es,
                              cudaMemcpyDeviceToHost, ctx->copy));
  }
  QL_CUDA(cudaEventRecord(hp->ev_alt_out, ctx->alt));
  QL_CUDA(cudaStreamWaitEvent(ctx->stream, hp->ev_alt_out, 0));
  // the compute stream joins the copy stream (the next apply must not overwrite c_dev before it has left), then the host waits
  QL_CUDA(cudaEventRecord(hp->ev_done, ctx->copy));
  QL_CUDA(cudaStreamWaitEvent(ctx->stream, hp->ev_done, 0));
  QL_CUDA(cudaStreamSynchronize(ctx->copy));
  ctx->launches = hp->launches;
  return QLB200_OK;
}

uint64_t qlb200_hostpipe_launches(const qlb200_hostpipe *hp) { return hp ? hp->launches : 0; }

// ---- accumulate form ------------------------------------------------------------------------------
struct qlb200_accum {
  AccumLayout L;
  int dtype = 0;
  double alpha[2] = {1.0, 0.0}, beta[2] = {0.0, 0.0};
};

int qlb200_accum_create(const qlb200_match *m, const qlb200_shell *c_old, int c_old_has_data, int allow_expand, int dtype,
                        const double *alpha2, const double *beta2, qlb200_accum **out) {
  if (!m || !out || !alpha2 || !beta2) return Fail(QLB200_ERR_ARG, "null argument");
  if (dtype != QLB200_F64 && dtype != QLB200_C64) return Fail(QLB200_ERR_ARG, "bad dtype");
  const double bi = dtype == QLB200_C64 ? beta2[1] : 0.0, ai = dtype == QLB200_C64 ? alpha2[1] : 0.0;
  const bool beta_zero = beta2[0] == 0.0 && bi == 0.0, beta_one = beta2[0] == 1.0 && bi == 0.0;
  // the reference's argument checks (contract_contiguous_axes.h:376-380, :683-688): no existing values to scale
  if (c_old == nullptr && !beta_zero) return Fail(QLB200_ERR_ARG, "accumulate into a default output requires beta == 0");
  if (c_old != nullptr && !c_old_has_data && !beta_zero && (c_old->nblk > 0 || c_old->rank == 0))
    return Fail(QLB200_ERR_ARG, "accumulate requires allocated output raw data unless beta == 0");
  qlb200_accum *a = new (std::nothrow) qlb200_accum();
  if (!a) return Fail(QLB200_ERR_NOMEM, "out of memory");
  a->dtype = dtype;
  a->alpha[0] = alpha2[0]; a->alpha[1] = ai; a->beta[0] = beta2[0]; a->beta[1] = bi;
  bool mismatch = false;
  std::string err = BuildAccumLayout(m->m, c_old, c_old_has_data != 0, allow_expand != 0, dtype, beta_zero, beta_one, &a->L, &mismatch);
  if (!err.empty()) { delete a; return Fail(mismatch ? QLB200_ERR_LAYOUT : QLB200_ERR_ARG, err); }
  *out = a;
  return QLB200_OK;
}
void qlb200_accum_destroy(qlb200_accum *a) { delete a; }
uint64_t qlb200_accum_nblk(const qlb200_accum *a) { return a ? a->L.blocks.size() : 0; }
uint64_t qlb200_accum_elems(const qlb200_accum *a) { return a ? a->L.elems : 0; }
int qlb200_accum_expanded(const qlb200_accum *a) { return a && a->L.expanded ? 1 : 0; }
int qlb200_accum_blocks(const qlb200_accum *a, uint64_t *blk_idx, uint32_t *blk_coors, uint32_t *shape, uint64_t *offset,
                        uint64_t *old_offset, uint8_t *touched) {
  if (!a) return Fail(QLB200_ERR_ARG, "null argument");
  const int r = a->L.rank;
  for (size_t b = 0; b < a->L.blocks.size(); ++b) {
    const CBlock &cb = a->L.blocks[b];
    if (blk_idx) blk_idx[b] = cb.blk_idx;
    if (offset) offset[b] = cb.offset;
    if (old_offset) old_offset[b] = a->L.old_off[b];
    if (touched) touched[b] = a->L.touched[b];
    for (int i = 0; i < r; ++i) {
      if (blk_coors) blk_coors[b * r + i] = cb.coors[i];
      if (shape) shape[b * r + i] = cb.shape[i];
    }
  }
  return QLB200_OK;
}
int qlb200_accum_get_stats(const qlb200_accum *a, qlb200_accum_stats *out) {
  if (!a || !out) return Fail(QLB200_ERR_ARG, "null argument");
  *out = a->L.stats;
  return QLB200_OK;
}

int qlb200_plan_create_accum(qlb200_ctx *ctx, const qlb200_match *m, const qlb200_accum *a, int dtype, uint32_t flags,
                             qlb200_plan **out) {
  if (!m || !a || !out) return Fail(QLB200_ERR_ARG, "null argument");
  if (dtype != a->dtype) return Fail(QLB200_ERR_ARG, "dtype differs from the accumulate layout's");
  const Match &mm = m->m;
  const AccumLayout &L = a->L;
  const bool beta_zero = a->beta[0] == 0.0 && a->beta[1] == 0.0, beta_one = a->beta[0] == 1.0 && a->beta[1] == 0.0;
  // tasks addressed in the RESULTING topology
  std::vector<qlb200_task> st = mm.SortedTasks();
  if (!mm.scalar) {
    for (qlb200_task &t : st) {
      const uint64_t u = L.req_to_union[t.c_ord];
      t.c_off = L.blocks[u].offset;
      t.c_ord = static_cast<uint32_t>(u);
    }
  }
  qlb200_plan *p = new (std::nothrow) qlb200_plan();
  if (!p) return Fail(QLB200_ERR_NOMEM, "out of memory");
  p->ctx = ctx;
  SetPlanKnobs(ctx, &flags, &p->h);
  std::vector<int32_t> ap(mm.a_perm.begin(), mm.a_perm.end()), bp(mm.b_perm.begin(), mm.b_perm.end());
  std::string err = BuildPlanHost(dtype, flags, static_cast<int>(mm.a_ctrct.size()), mm.a.rank, ap.data(), mm.a.nblk,
                                  mm.a.shape.data(), mm.a.offset.data(), mm.a.elems, mm.b.rank, bp.data(), mm.b.nblk,
                                  mm.b.shape.data(), mm.b.offset.data(), mm.b.elems, st, mm.scalar ? 1 : L.elems, &p->h);
  if (!err.empty()) { delete p; return Fail(QLB200_ERR_UNSUPPORTED, err); }
  p->accum = true;
  p->alpha[0] = a->alpha[0]; p->alpha[1] = a->alpha[1]; p->beta[0] = a->beta[0]; p->beta[1] = a->beta[1];
  p->c_old_elems = L.old_elems;
  p->acc_expanded = L.expanded;
  // every output block the contraction touches: where its old values lie (if any are to be read)
  auto set_in = [&](GemmGroup &g) {
    if (mm.scalar) { g.c_in_off = 0; g.beta_on = (!beta_zero && L.old_elems == 1) ? 1u : 0u; return; }
    auto it = std::lower_bound(L.blocks.begin(), L.blocks.end(), g.c_off, [](const CBlock &b, uint64_t off) { return b.offset < off; });
    const size_t u = size_t(it - L.blocks.begin());
    const bool is_new = L.c_default || L.old_off[u] == ~0ull;
    g.c_in_off = is_new ? 0 : L.old_off[u];
    g.beta_on = (!is_new && !beta_zero) ? 1u : 0u;
  };
  for (GemmGroup &g : p->h.groups) set_in(g);
  for (GemmGroup &g : p->h.part_groups) set_in(g);
  // output blocks the contraction does not touch: scaled in place, or scale-copied to their new place after an expansion
  if (mm.scalar) {
    if (st.empty() && L.old_elems == 1 && !beta_one) { p->acc_ranges.insert(p->acc_ranges.end(), {0ull, 0ull, 1ull}); }
  } else if (!L.c_default) {
    for (size_t u = 0; u < L.blocks.size(); ++u) {
      if (L.touched[u] || L.old_off[u] == ~0ull) continue;
      if (!L.expanded && beta_one) continue;
      p->acc_ranges.insert(p->acc_ranges.end(), {(unsigned long long) L.old_off[u], (unsigned long long) L.blocks[u].offset, (unsigned long long) L.blocks[u].size});
    }
  }
  if (ctx != nullptr) {
    int rc = FinishPlan(ctx, p);
    if (rc == QLB200_OK && !p->acc_ranges.empty()) {
      cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p->d_acc_ranges), p->acc_ranges.size() * sizeof(unsigned long long));
      if (e == cudaSuccess) e = cudaMemcpy(p->d_acc_ranges, p->acc_ranges.data(), p->acc_ranges.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) rc = Fail(QLB200_ERR_CUDA, CudaErr("accumulate range table", e));
    }
    if (rc != QLB200_OK) { p->d.Free(); if (p->d_acc_ranges) cudaFree(p->d_acc_ranges); delete p; return rc; }
  }
  *out = p;
  return QLB200_OK;
}

int qlb200_execute_accum(qlb200_ctx *ctx, qlb200_plan *p, const void *A, const void *B, const void *C_old, void *C_new, int mem_kind) {
  int ok = CheckExec(ctx, p, A, B);
  if (ok != QLB200_OK) return ok;
  if (!p->accum) return Fail(QLB200_ERR_ARG, "not an accumulate plan (qlb200_plan_create_accum)");
  if (!C_new) return Fail(QLB200_ERR_ARG, "null output");
  const bool beta_zero = p->beta[0] == 0.0 && p->beta[1] == 0.0;
  const bool need_old = !beta_zero && p->c_old_elems > 0;
  if (need_old && !C_old) return Fail(QLB200_ERR_ARG, "beta != 0 needs the existing output values");
  QL_CUDA(cudaSetDevice(ctx->device));
  const size_t es = ElemSize(p->h.dtype);
  const void *dA = A, *dB = B, *dCo = C_old;
  void *dCn = C_new;
  if (mem_kind == QLB200_MEM_HOST) {
    const size_t ab = Align256(p->h.a_elems * es), bb = Align256(p->h.b_elems * es), cn = Align256(p->h.c_elems * es),
                 co = need_old ? Align256(p->c_old_elems * es) : 0;
    int rc = EnsureArena(ctx, &ctx->stage, &ctx->stage_bytes, ab + bb + cn + co);
    if (rc != QLB200_OK) return rc;
    char *base = static_cast<char *>(ctx->stage);
    QL_CUDA(cudaMemcpyAsync(base, A, p->h.a_elems * es, cudaMemcpyHostToDevice, ctx->stream));
    QL_CUDA(cudaMemcpyAsync(base + ab, B, p->h.b_elems * es, cudaMemcpyHostToDevice, ctx->stream));
    // same topology: the old values are staged where the result is built (in place: untouched blocks with beta == 1 stay put)
    char *old_at = p->acc_expanded ? base + ab + bb + cn : base + ab + bb;
    if (need_old) QL_CUDA(cudaMemcpyAsync(old_at, C_old, p->c_old_elems * es, cudaMemcpyHostToDevice, ctx->stream));
    dA = base; dB = base + ab; dCn = base + ab + bb; dCo = need_old ? old_at : nullptr;
  } else if (mem_kind != QLB200_MEM_DEVICE) {
    return Fail(QLB200_ERR_ARG, "bad mem_kind");
  } else if (need_old && !p->acc_expanded && C_old != C_new) {
    return Fail(QLB200_ERR_ARG, "same output topology: the accumulate runs in place, pass C_old == C_new");
  } else if (need_old && p->acc_expanded && C_old == C_new) {
    return Fail(QLB200_ERR_ARG, "expanded output topology: C_new must be a different buffer than C_old");
  }
  int rc = qlb200_execute_permute(ctx, p, dA, dB);
  if (rc != QLB200_OK) return rc;
  uint64_t launched = ctx->launches;
  if (!p->acc_ranges.empty()) {
    // beta == 0: the kernel stores zeros without reading, so dCo may be null
    QL_CUDA(LaunchScaleCopyRanges(p->h.dtype, p->d_acc_ranges, static_cast<uint32_t>(p->acc_ranges.size() / 3), dCo ? dCo : dCn, dCn,
                                  p->beta[0], p->beta[1], ctx->num_sms, ctx->stream));
    ++launched; ++ctx->total_launches;
  }
  void *outs[1] = {dCn};
  rc = ExecuteGemm(ctx, p, dA, dB, outs, 1, 0, dCo);
  if (rc != QLB200_OK) return rc;
  ctx->launches += launched;
  if (mem_kind == QLB200_MEM_HOST) {
    QL_CUDA(cudaMemcpyAsync(C_new, dCn, p->h.c_elems * es, cudaMemcpyDeviceToHost, ctx->stream));
    QL_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return QLB200_OK;
}

// ---- matrix-free axis operations -----------------------------------------------------------------
struct qlb200_axis { AxisMatch m; };

int qlb200_axis_create(const qlb200_shell *in, int32_t nops, const qlb200_shell *op1, int32_t axis1, const qlb200_shell *op2,
                       int32_t axis2, qlb200_axis **out) {
  if (!in || !op1 || !out || (nops == 2 && !op2)) return Fail(QLB200_ERR_ARG, "null argument");
  qlb200_axis *a = new (std::nothrow) qlb200_axis();
  if (!a) return Fail(QLB200_ERR_NOMEM, "out of memory");
  std::string err = BuildAxisMatch(in, nops, op1, axis1, op2, axis2, &a->m);
  if (!err.empty()) { delete a; return Fail(QLB200_ERR_ARG, err); }
  *out = a;
  return QLB200_OK;
}
void qlb200_axis_destroy(qlb200_axis *a) { delete a; }
uint64_t qlb200_axis_out_nblk(const qlb200_axis *a) { return a ? a->m.out_blocks.size() : 0; }
uint64_t qlb200_axis_out_elems(const qlb200_axis *a) { return a ? a->m.out_elems : 0; }
uint64_t qlb200_axis_nterm(const qlb200_axis *a) { return a ? a->m.terms.size() : 0; }
int qlb200_axis_out_blocks(const qlb200_axis *a, uint64_t *blk_idx, uint32_t *blk_coors, uint32_t *shape, uint64_t *offset) {
  if (!a) return Fail(QLB200_ERR_ARG, "null argument");
  const int r = a->m.in.rank;
  for (size_t b = 0; b < a->m.out_blocks.size(); ++b) {
    const CBlock &cb = a->m.out_blocks[b];
    if (blk_idx) blk_idx[b] = cb.blk_idx;
    if (offset) offset[b] = cb.offset;
    for (int i = 0; i < r; ++i) {
      if (blk_coors) blk_coors[b * r + i] = cb.coors[i];
      if (shape) shape[b * r + i] = cb.shape[i];
    }
  }
  return QLB200_OK;
}

struct qlb200_axis_plan {
  qlb200_ctx *ctx = nullptr;
  int dtype = 0, nops = 1;
  bool swapped = false;                 // the kernel's first operator acts on the EARLIER axis: op1 / op2 swapped if axis1 > axis2
  uint64_t in_elems = 0, op_elems[2] = {0, 0}, out_elems = 0;
  uint64_t read_bytes = 0, write_bytes = 0;
  std::vector<AxisOut> outs;
  std::vector<AxisFlatTerm> terms;
  std::vector<AxisItem> items;
  AxisOut *d_outs = nullptr;
  AxisFlatTerm *d_terms = nullptr;
  AxisItem *d_items = nullptr;
};

void qlb200_axis_plan_destroy(qlb200_axis_plan *p) {
  if (!p) return;
  if (p->ctx) { cudaSetDevice(p->ctx->device); cudaStreamSynchronize(p->ctx->stream); }
  cudaFree(p->d_outs); cudaFree(p->d_terms); cudaFree(p->d_items);
  delete p;
}

int qlb200_axis_plan_create(qlb200_ctx *ctx, const qlb200_axis *a, int dtype, qlb200_axis_plan **out) {
  if (!ctx || !a || !out) return Fail(QLB200_ERR_ARG, "null argument");
  if (dtype != QLB200_F64 && dtype != QLB200_C64) return Fail(QLB200_ERR_ARG, "bad dtype");
  const AxisMatch &m = a->m;
  const int r = m.in.rank;
  qlb200_axis_plan *p = new (std::nothrow) qlb200_axis_plan();
  if (!p) return Fail(QLB200_ERR_NOMEM, "out of memory");
  p->ctx = ctx; p->dtype = dtype; p->nops = m.nops;
  p->in_elems = m.in.elems; p->op_elems[0] = m.op[0].elems; p->op_elems[1] = m.nops == 2 ? m.op[1].elems : 0; p->out_elems = m.out_elems;
  // kernel operator A acts on the earlier axis
  int oa = 0, ob = 1;
  if (m.nops == 2 && m.axis[0] > m.axis[1]) { oa = 1; ob = 0; p->swapped = true; }
  const int a1 = m.axis[oa], a2 = m.nops == 2 ? m.axis[ob] : -1;
  const size_t es = ElemSize(dtype);
  for (size_t b = 0; b < m.out_blocks.size(); ++b) {
    const CBlock &cb = m.out_blocks[b];
    AxisOut o;
    std::memset(&o, 0, sizeof(o));
    o.out_off = cb.offset; o.size = uint32_t(cb.size);
    uint64_t P1 = 1, P2 = 1;
    if (a2 >= 0) {
      for (int i = a1 + 1; i < a2; ++i) P1 *= cb.shape[i];
      for (int i = a2 + 1; i < r; ++i) P2 *= cb.shape[i];
    } else {
      for (int i = a1 + 1; i < r; ++i) P2 *= cb.shape[i];
    }
    o.J1 = cb.shape[a1]; o.P1 = uint32_t(P1); o.J2 = a2 >= 0 ? cb.shape[a2] : 1u; o.P2 = uint32_t(P2);
    o.term_begin = uint32_t(p->terms.size());
    if (o.J1 > uint32_t(kAxisMaxDim) || o.J2 > uint32_t(kAxisMaxDim)) { delete p; return Fail(QLB200_ERR_UNSUPPORTED, "operator block wider than 8: use the contraction path"); }
    for (uint32_t t = m.term_begin[b]; t < m.term_begin[b + 1]; ++t) {
      const AxisTerm &tm = m.terms[t];
      const uint32_t opa_blk = oa == 0 ? tm.op1_ord : tm.op2_ord, opb_blk = oa == 0 ? tm.op2_ord : tm.op1_ord;
      const uint32_t I1 = m.in.shape[uint64_t(tm.in_ord) * r + a1], I2 = a2 >= 0 ? m.in.shape[uint64_t(tm.in_ord) * r + a2] : 1u;
      if (I1 > uint32_t(kAxisMaxDim) || I2 > uint32_t(kAxisMaxDim)) { delete p; return Fail(QLB200_ERR_UNSUPPORTED, "operator block taller than 8: use the contraction path"); }
      const uint64_t s_i1 = P1 * I2 * P2, s_i2 = P2;
      if (uint64_t(I1) * s_i1 >= (1ull << 32)) { delete p; return Fail(QLB200_ERR_UNSUPPORTED, "input block too large"); }
      for (uint32_t i1 = 0; i1 < I1; ++i1)
        for (uint32_t i2 = 0; i2 < I2; ++i2) {
          AxisFlatTerm ft;
          ft.in_base = m.in.offset[tm.in_ord] + i1 * s_i1 + i2 * s_i2;
          ft.c1_off = m.op[oa].offset[opa_blk] + uint64_t(i1) * o.J1;
          ft.c2_off = a2 >= 0 ? m.op[ob].offset[opb_blk] + uint64_t(i2) * o.J2 : 0;
          ft.sp0 = uint32_t(I1 * s_i1); ft.sp1 = uint32_t(I2 * P2);
          p->terms.push_back(ft);
        }
      p->read_bytes += m.in.size[tm.in_ord] * es;
    }
    o.term_end = uint32_t(p->terms.size());
    p->outs.push_back(o);
    for (uint64_t e = 0; e < cb.size; e += 1024) p->items.push_back({uint32_t(b), uint32_t(e)});
    p->write_bytes += cb.size * es;
  }
  if (p->items.size() >= (1ull << 32)) { delete p; return Fail(QLB200_ERR_UNSUPPORTED, "too many work items"); }
  QL_CUDA(cudaSetDevice(ctx->device));
  int rc = Upload(p->outs, &p->d_outs, ctx->stream);
  if (rc == QLB200_OK) rc = Upload(p->terms, &p->d_terms, ctx->stream);
  if (rc == QLB200_OK) rc = Upload(p->items, &p->d_items, ctx->stream);
  if (rc != QLB200_OK) { qlb200_axis_plan_destroy(p); return rc; }
  QL_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = p;
  return QLB200_OK;
}

int qlb200_axis_plan_bytes(const qlb200_axis_plan *p, uint64_t *read_bytes, uint64_t *write_bytes) {
  if (!p) return Fail(QLB200_ERR_ARG, "null argument");
  if (read_bytes) *read_bytes = p->read_bytes;
  if (write_bytes) *write_bytes = p->write_bytes;
  return QLB200_OK;
}

int qlb200_axis_execute(qlb200_ctx *ctx, qlb200_axis_plan *p, const void *in, const void *op1, const void *op2, void *out, int mem_kind) {
  if (!ctx || !p || !in || !op1 || !out || (p->nops == 2 && !op2)) return Fail(QLB200_ERR_ARG, "null argument");
  if (p->ctx != ctx) return Fail(QLB200_ERR_ARG, "plan belongs to another context");
  QL_CUDA(cudaSetDevice(ctx->device));
  const size_t es = ElemSize(p->dtype);
  const void *d_in = in, *d_o1 = op1, *d_o2 = p->nops == 2 ? op2 : nullptr;
  void *d_out = out;
  if (mem_kind == QLB200_MEM_HOST) {
    const size_t ib = Align256(p->in_elems * es), o1b = Align256(p->op_elems[0] * es), o2b = Align256(p->op_elems[1] * es), ob = Align256(p->out_elems * es);
    int rc = EnsureArena(ctx, &ctx->stage, &ctx->stage_bytes, ib + o1b + o2b + ob);
    if (rc != QLB200_OK) return rc;
    char *base = static_cast<char *>(ctx->stage);
    QL_CUDA(cudaMemcpyAsync(base, in, p->in_elems * es, cudaMemcpyHostToDevice, ctx->stream));
    QL_CUDA(cudaMemcpyAsync(base + ib, op1, p->op_elems[0] * es, cudaMemcpyHostToDevice, ctx->stream));
    if (p->nops == 2) QL_CUDA(cudaMemcpyAsync(base + ib + o1b, op2, p->op_elems[1] * es, cudaMemcpyHostToDevice, ctx->stream));
    d_in = base; d_o1 = base + ib; d_o2 = p->nops == 2 ? base + ib + o1b : nullptr; d_out = base + ib + o1b + o2b;
  } else if (mem_kind != QLB200_MEM_DEVICE) {
    return Fail(QLB200_ERR_ARG, "bad mem_kind");
  }
  if (p->swapped) std::swap(d_o1, d_o2);
  ctx->launches = 0;
  if (!p->items.empty()) {
    QL_CUDA(LaunchAxisApply(p->dtype, p->d_outs, p->d_terms, p->d_items, static_cast<uint32_t>(p->items.size()), d_in, d_o1, d_o2, d_out,
                            ctx->num_sms, ctx->stream));
    ctx->launches = 1; ++ctx->total_launches;
  }
  if (mem_kind == QLB200_MEM_HOST) {
    QL_CUDA(cudaMemcpyAsync(out, d_out, p->out_elems * es, cudaMemcpyDeviceToHost, ctx->stream));
    QL_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return QLB200_OK;
}

// ---- whole-tensor transpose ---------------------------------------------------------------------
int qlb200_tplan_create(qlb200_ctx *ctx, const qlb200_shell *t, const int32_t *perm, int dtype, qlb200_tplan **out) {
  if (!ctx || !t || !perm || !out) return Fail(QLB200_ERR_ARG, "null argument");
  if (dtype != QLB200_F64 && dtype != QLB200_C64) return Fail(QLB200_ERR_ARG, "bad dtype");
  Shell s;
  std::string err = s.Load(t);
  if (!err.empty()) return Fail(QLB200_ERR_ARG, err);
  const int r = s.rank;
  std::vector<char> seen(r, 0);
  for (int i = 0; i < r; ++i) {
    if (perm[i] < 0 || perm[i] >= r || seen[perm[i]]) return Fail(QLB200_ERR_ARG, "perm is not a permutation");
    seen[perm[i]] = 1;
  }
  qlb200_tplan *p = new (std::nothrow) qlb200_tplan();
  if (!p) return Fail(QLB200_ERR_NOMEM, "out of memory");
  p->ctx = ctx; p->dtype = dtype; p->elems = s.elems; p->rank = r;
  // transposed blocks: new coordinates/shape, new blk_idx over the permuted sector counts
  struct TB { uint64_t idx; uint64_t src; };
  std::vector<TB> tb(s.nblk);
  std::vector<uint32_t> nsct_t(r);
  for (int i = 0; i < r; ++i) nsct_t[i] = s.nsct[perm[i]];
  for (uint64_t b = 0; b < s.nblk; ++b) {
    uint64_t idx = 0;
    for (int i = 0; i < r; ++i) idx = idx * nsct_t[i] + s.coors[b * r + perm[i]];
    tb[b] = {idx, b};
  }
  std::sort(tb.begin(), tb.end(), [](const TB &x, const TB &y) { return x.idx < y.idx; });
  p->blk_idx.resize(s.nblk); p->offset.resize(s.nblk); p->scale.resize(s.nblk);
  p->coors.resize(s.nblk * r); p->shape.resize(s.nblk * r);
  uint64_t off = 0;
  p->perm_tile_base.push_back(0);
  uint8_t par[QLB200_MAX_RANK];
  for (uint64_t i = 0; i < s.nblk; ++i) {
    const uint64_t b = tb[i].src;
    p->blk_idx[i] = tb[i].idx; p->offset[i] = off;
    for (int a = 0; a < r; ++a) { p->coors[i * r + a] = s.coors[b * r + perm[a]]; p->shape[i * r + a] = s.shape[b * r + perm[a]]; }
    int sign = 1;
    if (s.fermionic()) {
      for (int a = 0; a < r; ++a) par[a] = s.parity[s.sct_base[a] + s.coors[b * r + a]];
      sign = FermionReorderSign(par, r, perm);
    }
    p->scale[i] = static_cast<int8_t>(sign);
    if (s.size[b] >= (1ull << 32)) { delete p; return Fail(QLB200_ERR_UNSUPPORTED, "block with 2^32 or more elements"); }
    uint64_t nt = 0;
    PermBlk d = MakePermBlk(r, &s.shape[b * r], perm, s.offset[b], off, 0, float(sign), &nt, int(ElemSize(dtype)));
    if (p->perm_tile_base.back() + nt >= (1ull << 32)) { delete p; return Fail(QLB200_ERR_UNSUPPORTED, "too many permute tiles"); }
    p->perm_blks.push_back(d);
    p->perm_tile_base.push_back(static_cast<uint32_t>(p->perm_tile_base.back() + nt));
    off += s.size[b];
  }
  QL_CUDA(cudaSetDevice(ctx->device));
  int rc = Upload(p->perm_blks, &p->d.perm_blks, ctx->stream);
  if (rc == QLB200_OK) rc = Upload(p->perm_tile_base, &p->d.perm_tile_base, ctx->stream);
  if (rc != QLB200_OK) { p->d.Free(); delete p; return rc; }
  QL_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = p;
  return QLB200_OK;
}
void qlb200_tplan_destroy(qlb200_tplan *p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  p->d.Free();
  delete p;
}
uint64_t qlb200_tplan_nblk(const qlb200_tplan *p) { return p->blk_idx.size(); }
int qlb200_tplan_blocks(const qlb200_tplan *p, uint64_t *blk_idx, uint32_t *blk_coors, uint32_t *shape,
                        uint64_t *offset, int8_t *scale) {
  const size_t n = p->blk_idx.size();
  if (blk_idx) std::memcpy(blk_idx, p->blk_idx.data(), n * sizeof(uint64_t));
  if (offset) std::memcpy(offset, p->offset.data(), n * sizeof(uint64_t));
  if (blk_coors) std::memcpy(blk_coors, p->coors.data(), n * p->rank * sizeof(uint32_t));
  if (shape) std::memcpy(shape, p->shape.data(), n * p->rank * sizeof(uint32_t));
  if (scale) std::memcpy(scale, p->scale.data(), n * sizeof(int8_t));
  return QLB200_OK;
}
int qlb200_transpose_execute(qlb200_ctx *ctx, qlb200_tplan *p, const void *src, void *dst, int mem_kind) {
  if (!ctx || !p || !src || !dst) return Fail(QLB200_ERR_ARG, "null argument");
  QL_CUDA(cudaSetDevice(ctx->device));
  const size_t es = ElemSize(p->dtype);
  const void *ds = src; void *dd = dst;
  const size_t bytes = p->elems * es;
  if (mem_kind == QLB200_MEM_HOST) {
    int rc = EnsureArena(ctx, &ctx->stage, &ctx->stage_bytes, 2 * Align256(bytes));
    if (rc != QLB200_OK) return rc;
    char *base = static_cast<char *>(ctx->stage);
    QL_CUDA(cudaMemcpyAsync(base, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ds = base; dd = base + Align256(bytes);
  }
  ctx->launches = 0;
  const uint32_t ntiles = p->perm_tile_base.back();
  if (ntiles > 0) {
    QL_CUDA(LaunchPermute(p->dtype, p->d.perm_blks, p->d.perm_tile_base, static_cast<uint32_t>(p->perm_blks.size()), ntiles,
                          ds, ds, dd, dd, ctx->num_sms, ctx->stream));
    ctx->launches += 1; ctx->total_launches += 1;
  }
  if (mem_kind == QLB200_MEM_HOST) {
    QL_CUDA(cudaMemcpyAsync(dst, dd, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    QL_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return QLB200_OK;
}

// ---- batched range copy ---------------------------------------------------------------------------
int qlb200_cplan_create(qlb200_ctx *ctx, int dtype, uint64_t n, const uint64_t *src_off, const uint64_t *dst_off,
                        const uint64_t *len, qlb200_tplan **out) {
  if (!ctx || !out || (n && (!src_off || !dst_off || !len))) return Fail(QLB200_ERR_ARG, "null argument");
  if (dtype != QLB200_F64 && dtype != QLB200_C64) return Fail(QLB200_ERR_ARG, "bad dtype");
  qlb200_tplan *p = new (std::nothrow) qlb200_tplan();
  if (!p) return Fail(QLB200_ERR_NOMEM, "out of memory");
  p->ctx = ctx; p->dtype = dtype; p->rank = 1;
  p->perm_tile_base.push_back(0);
  const int32_t ident[1] = {0};
  for (uint64_t i = 0; i < n; ++i) {
    // ranges longer than 2^31 elements are split so extents stay 32-bit
    for (uint64_t done = 0; done < len[i];) {
      const uint64_t chunk = std::min<uint64_t>(len[i] - done, 1ull << 31);
      const uint32_t shape[1] = {static_cast<uint32_t>(chunk)};
      uint64_t nt = 0;
      PermBlk d = MakePermBlk(1, shape, ident, src_off[i] + done, dst_off[i] + done, 0, 1.0f, &nt);
      if (p->perm_tile_base.back() + nt >= (1ull << 32)) { delete p; return Fail(QLB200_ERR_UNSUPPORTED, "too many copy tiles"); }
      p->perm_blks.push_back(d);
      p->perm_tile_base.push_back(static_cast<uint32_t>(p->perm_tile_base.back() + nt));
      p->elems += chunk;
      done += chunk;
    }
  }
  QL_CUDA(cudaSetDevice(ctx->device));
  int rc = Upload(p->perm_blks, &p->d.perm_blks, ctx->stream);
  if (rc == QLB200_OK) rc = Upload(p->perm_tile_base, &p->d.perm_tile_base, ctx->stream);
  if (rc != QLB200_OK) { p->d.Free(); delete p; return rc; }
  QL_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = p;
  return QLB200_OK;
}
int qlb200_copy_execute(qlb200_ctx *ctx, qlb200_tplan *p, const void *src, void *dst) {
  if (!ctx || !p || !src || !dst) return Fail(QLB200_ERR_ARG, "null argument");
  QL_CUDA(cudaSetDevice(ctx->device));
  ctx->launches = 0;
  const uint32_t ntiles = p->perm_tile_base.back();
  if (ntiles > 0) {
    QL_CUDA(LaunchPermute(p->dtype, p->d.perm_blks, p->d.perm_tile_base, static_cast<uint32_t>(p->perm_blks.size()), ntiles,
                          src, src, dst, dst, ctx->num_sms, ctx->stream));
    ctx->launches += 1; ctx->total_launches += 1;
  }
  return QLB200_OK;
}

}  // extern "C"
