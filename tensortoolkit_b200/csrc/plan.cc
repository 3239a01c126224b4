// plan.cc -- host-side construction of the device descriptor tables (permute blocks, GEMM groups,
// tile lists, multi-GPU row partition).  No reference counterpart: the reference walks tasks one
// by one (global_operations.h:919-982); here the whole contraction is flattened into tables once.
#include "plan.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <cstring>
#include <functional>
#include <numeric>

namespace qlb200 {

namespace {
uint32_t Pow2Ceil(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }
uint32_t Log2(uint32_t v) { uint32_t l = 0; while ((1u << l) < v) ++l; return l; }
constexpr uint32_t kSmemCap = 2304;
}  // namespace

PermBlk MakePermBlk(int rank, const uint32_t *shape, const int32_t *perm, uint64_t src_off, uint64_t dst_off,
                    uint32_t src_sel, float scale, uint64_t *ntiles_out, int elem_bytes) {
  PermBlk d;
  std::memset(&d, 0, sizeof(d));
  d.src_off = src_off; d.dst_off = dst_off; d.src_sel = src_sel; d.scale = scale;
  // input strides (row-major)
  uint64_t istr[QLB200_MAX_RANK];
  uint64_t s = 1;
  for (int i = rank - 1; i >= 0; --i) { istr[i] = s; s *= shape[i]; }
  // output-ordered axes, size-1 axes dropped, mergeable neighbours fused
  uint64_t ext[QLB200_MAX_RANK], sst[QLB200_MAX_RANK];
  int nd = 0;
  for (int j = 0; j < rank; ++j) {
    const uint64_t e = shape[perm[j]], st = istr[perm[j]];
    if (e == 1) continue;
    if (nd > 0 && sst[nd - 1] == st * e) { ext[nd - 1] *= e; sst[nd - 1] = st; }
    else { ext[nd] = e; sst[nd] = st; ++nd; }
  }
  if (nd == 0) { ext[0] = 1; sst[0] = 1; nd = 1; }
  d.nd = nd;
  uint64_t ds = 1;
  for (int j = nd - 1; j >= 0; --j) {
    d.ext[j] = static_cast<uint32_t>(ext[j]); d.sstr[j] = static_cast<uint32_t>(sst[j]);
    d.dstr[j] = static_cast<uint32_t>(ds); ds *= ext[j];
  }
  uint32_t jin = nd - 1;
  for (int j = 0; j < nd; ++j) if (sst[j] == 1) jin = j;
  d.jin = jin;
  constexpr uint32_t kVecMaxRun = 64;       // runs shorter than this are moved in groups (run mode)
  if (jin == uint32_t(nd - 1) && nd >= 3 && d.ext[nd - 1] < kVecMaxRun) {
    // Short runs shared by source and destination, e.g. (n1, k, n2) -> (k, n1, n2) with small n2: a tile of single runs
    // would move a few hundred bytes per CTA iteration.  Transpose RUNS instead: js = the axis that follows the run in
    // the source (stride V), jd = nd-2 follows it in the destination (stride V); a TI x TO tile of runs is read as TO
    // contiguous pieces of TI*V elements and written as TI contiguous pieces of TO*V elements.
    const uint32_t V = d.ext[nd - 1];
    uint32_t js = 0;
    for (int j = 0; j < nd - 1; ++j) if (sst[j] < sst[js]) js = uint32_t(j);
    const uint32_t jd = uint32_t(nd - 2);
    // js != jd: otherwise the two axes would have been merged above (dense block: the stride that follows V is V itself)
    if (js != jd && sst[js] == V) {
      d.vec = V; d.jin = js; d.jout = jd;
      // bulk-copy path: every run starts 16-byte aligned in the source and in the destination and is a multiple of 16
      // bytes long (all other strides of a dense block are multiples of V), and nothing is scaled on the way
      const uint64_t eb = uint64_t(elem_bytes);
      d.bulk = (eb != 0 && scale == 1.0f && (V * eb) % 16 == 0 && (src_off * eb) % 16 == 0 && (dst_off * eb) % 16 == 0) ? 1u : 0u;
      const uint32_t pad = d.bulk ? 0u : 1u;            // the element-wise path pads rows against bank conflicts
      const uint32_t ei = d.ext[js], eo = d.ext[jd];
      uint32_t side = 1;
      while ((side + 1) * (side + 1) * V + (side + 1) <= kSmemCap) ++side;      // square-ish tile of runs
      uint32_t TI = std::min(ei, side), TO = std::min(eo, side);
      // a short axis leaves room for more of the other one
      if (TI < side) TO = std::min<uint32_t>(eo, kSmemCap / (TI * V + pad));
      else if (TO < side) TI = std::min<uint32_t>(ei, (kSmemCap / TO - pad) / V);
      while ((TI * V + pad) * TO > kSmemCap) { if (TO > 1) --TO; else --TI; }
      d.TI = TI; d.TO = TO;
      d.nti = (ei + TI - 1) / TI; d.nto = (eo + TO - 1) / TO;
      d.txi_log2 = d.txo_log2 = 0;
      uint64_t nt = uint64_t(d.nti) * d.nto;
      for (int j = 0; j < nd - 1; ++j) if (uint32_t(j) != js && uint32_t(j) != jd) nt *= d.ext[j];
      *ntiles_out = nt;
      return d;
    }
  }
  if (jin == uint32_t(nd - 1)) {            // contiguous runs on both sides
    d.jout = nd >= 2 ? nd - 2 : jin;
    d.TI = std::min<uint32_t>(d.ext[jin], 2048);
    d.TO = d.jout != jin ? std::min<uint32_t>(d.ext[d.jout], std::max<uint32_t>(1, 2048 / d.TI)) : 1;
  } else {
    d.jout = nd - 1;
    const uint32_t ei = d.ext[jin], eo = d.ext[nd - 1];
    uint32_t TO = std::min<uint32_t>(eo, 32), TI = std::min<uint32_t>(ei, 64);
    if (TO < 32) {                          // short output runs: take more of the input axis
      uint32_t cap = kSmemCap / TO;         // (TI|1) <= cap
      uint32_t t = cap >= 1 ? ((cap & 1) ? cap - 1 : cap - 2) : 1;   // largest TI with (TI|1) <= cap
      if (t < 1) t = 1;
      TI = std::min<uint32_t>(ei, std::min<uint32_t>(std::max<uint32_t>(t, 1), 1024));
    }
    if (TI < 64) TO = std::min<uint32_t>(eo, std::min<uint32_t>(kSmemCap / (TI | 1u), 1024));
    while ((TI | 1u) * TO > kSmemCap) { if (TO > 1) --TO; else --TI; }
    d.TI = TI; d.TO = TO;
  }
  d.nti = (d.ext[jin] + d.TI - 1) / d.TI;
  d.nto = d.jout != jin ? (d.ext[d.jout] + d.TO - 1) / d.TO : 1;
  d.txi_log2 = Log2(std::min<uint32_t>(256, Pow2Ceil(d.TI)));
  d.txo_log2 = Log2(std::min<uint32_t>(256, Pow2Ceil(d.TO)));
  uint64_t nt = uint64_t(d.nti) * d.nto;
  for (int j = 0; j < nd; ++j) if (uint32_t(j) != d.jin && uint32_t(j) != d.jout) nt *= d.ext[j];
  *ntiles_out = nt;
  return d;
}

// ------------------------------------------------------------------------------------------------
// Operand layouts.  After the permutation an operand block is a row-major R x Cc matrix (A: m x k,
// B: k x n).  Many blocks do not need the permute kernel at all:
//   kBlkDirect  the canonicalised permutation is the identity (all moved axes have extent 1, as the
//               physical legs of a DMRG tensor block do) -> the GEMM reads the caller's buffer;
//   kBlkTrans   it is one 2-D transposition whose cut coincides with the row/column cut -> the block
//               is a row-major Cc x R matrix in the caller's buffer and the GEMM loads it transposed;
//   kBlkPermute everything else goes through the batched permute kernel into the workspace.
// ------------------------------------------------------------------------------------------------
//   kBlkView    (B only) three merged axes (k | n1, n2) whose last one is contiguous in the source, e.g. a block stored
//               (n1, k, n2): the k x (n1 n2) matrix is a strided VIEW of the stored block, rows made of n1 runs of n2
//               contiguous elements (GemmTask::b_rs / b_cs / b_run) -> read in place as well
enum : uint8_t { kBlkDirect = 0, kBlkTrans = 1, kBlkPermute = 2, kBlkView = 3 };

struct BlockView { uint32_t rs = 0, cs = 0, run = 0; };

static uint8_t ClassifyBlock(int rank, const uint32_t *shape, const int32_t *perm, uint64_t rows, bool allow_trans,
                             bool allow_view = false, BlockView *view = nullptr) {
  uint64_t istr[QLB200_MAX_RANK];
  uint64_t s = 1;
  for (int i = rank - 1; i >= 0; --i) { istr[i] = s; s *= shape[i]; }
  uint64_t ext[QLB200_MAX_RANK], sst[QLB200_MAX_RANK];
  int nd = 0;
  for (int j = 0; j < rank; ++j) {
    const uint64_t e = shape[perm[j]], st = istr[perm[j]];
    if (e == 1) continue;
    if (nd > 0 && sst[nd - 1] == st * e) { ext[nd - 1] *= e; sst[nd - 1] = st; }
    else { ext[nd] = e; sst[nd] = st; ++nd; }
  }
  if (nd <= 1) return kBlkDirect;
  if (nd == 2 && allow_trans && ext[0] == rows) return kBlkTrans;
  // runs shorter than 4 elements would make every copy its own 32-byte sector: leave those to the permute kernel
  if (nd == 3 && allow_view && view != nullptr && ext[0] == rows && sst[2] == 1 && ext[2] >= 4 && sst[0] < (1ull << 32) &&
      sst[1] < (1ull << 32)) {
    view->rs = uint32_t(sst[0]); view->cs = uint32_t(sst[1]); view->run = uint32_t(ext[2]);
    return kBlkView;
  }
  return kBlkPermute;
}

struct OperandPlan {
  std::vector<uint8_t> mode;      // per block (unused blocks: kBlkDirect)
  std::vector<BlockView> view;    // per block, for kBlkView
  uint64_t permute_elems = 0;
};

static OperandPlan ClassifyOperand(int rank, const int32_t *perm, uint64_t n, const uint32_t *shape,
                                   const std::vector<char> &used, const std::vector<uint64_t> &rows,
                                   bool tensor_trans, bool per_block, bool allow_trans, bool allow_view = false) {
  OperandPlan op;
  op.mode.assign(n, kBlkDirect);
  op.view.assign(allow_view ? n : 0, BlockView());
  if (!tensor_trans) return op;
  for (uint64_t b = 0; b < n; ++b) {
    if (!used[b]) continue;
    op.mode[b] = per_block ? ClassifyBlock(rank, shape + b * rank, perm, rows[b], allow_trans, allow_view, allow_view ? &op.view[b] : nullptr)
                           : uint8_t(kBlkPermute);
    if (op.mode[b] == kBlkPermute) {
      uint64_t sz = 1;
      for (int i = 0; i < rank; ++i) sz *= shape[b * rank + i];
      op.permute_elems += sz;
    }
  }
  return op;
}

static std::string AddPermBlocks(PlanHost *h, int rank, const int32_t *perm, const uint32_t *shape,
                                 const uint64_t *off, const std::vector<uint8_t> &mode, const std::vector<char> &used,
                                 uint32_t src_sel, std::vector<uint64_t> *new_off, uint64_t *ws_elems, uint64_t *moved) {
  const uint64_t n = used.size();
  new_off->assign(n, 0);
  uint64_t ws = 0;
  for (uint64_t b = 0; b < n; ++b) {
    if (!used[b] || mode[b] != kBlkPermute) continue;
    uint64_t sz = 1;
    for (int i = 0; i < rank; ++i) sz *= shape[b * rank + i];
    if (sz >= (1ull << 32)) return "block with 2^32 or more elements";
    (*new_off)[b] = ws;
    uint64_t nt = 0;
    PermBlk d = MakePermBlk(rank, shape + b * rank, perm, off[b], ws, src_sel, 1.0f, &nt, h->dtype == QLB200_C64 ? 16 : 8);
    const uint64_t base = h->perm_tile_base.empty() ? 0 : h->perm_tile_base.back();
    if (h->perm_tile_base.empty()) h->perm_tile_base.push_back(0);
    if (base + nt >= (1ull << 32)) return "too many permute tiles";
    h->perm_blks.push_back(d);
    h->perm_tile_base.push_back(static_cast<uint32_t>(h->perm_tile_base.back() + nt));
    h->perm_blks_all.push_back(d);
    h->perm_ntiles_all.push_back(static_cast<uint32_t>(nt));
    h->perm_owner_all.push_back((uint64_t(src_sel) << 63) | b);
    h->perm_size_all.push_back(sz);
    ws += (sz + 1) & ~1ull;        // keep every permuted block 16-byte aligned for doubles
    *moved += sz;
  }
  *ws_elems = ws;
  return "";
}

std::string BuildPlanHost(int dtype, uint32_t flags, int nctrct, int a_rank, const int32_t *a_perm_in,
                          uint64_t na, const uint32_t *a_shape, const uint64_t *a_off, uint64_t a_elems,
                          int b_rank, const int32_t *b_perm_in, uint64_t nb, const uint32_t *b_shape,
                          const uint64_t *b_off, uint64_t b_elems, const std::vector<qlb200_task> &st,
                          uint64_t c_elems, PlanHost *h) {
  h->dtype = dtype; h->flags = flags;
  h->a_elems = a_elems; h->b_elems = b_elems; h->c_elems = c_elems;
  std::vector<char> a_used(na, 0), b_used(nb, 0);
  std::vector<uint64_t> a_rows(na, 0), b_rows(nb, 0);   // row count of the permuted block: m for A, k for B
  for (const auto &t : st) {
    if (t.a_ord >= na || t.b_ord >= nb) return "task references a block ordinal out of range";
    a_used[t.a_ord] = 1; b_used[t.b_ord] = 1;
    a_rows[t.a_ord] = t.m; b_rows[t.b_ord] = t.k;
  }
  auto is_ident = [](int rank, const int32_t *perm) {
    if (!perm) return true;
    for (int i = 0; i < rank; ++i) if (perm[i] != i) return false;
    return true;
  };
  std::vector<int32_t> a_perm(a_rank), b_perm(b_rank);
  for (int i = 0; i < a_rank; ++i) a_perm[i] = a_perm_in ? a_perm_in[i] : i;
  for (int i = 0; i < b_rank; ++i) b_perm[i] = b_perm_in ? b_perm_in[i] : i;

  // The GEMM kernels read blocks in place (direct or 2-D transposed) unless the caller forces the permute pass.
  const bool per_block = !(flags & QLB200_PLAN_PERMUTE_ALL);
  const bool allow_trans = per_block;

  // The order of the contracted axes inside the k index is free as long as A and B agree (it only
  // permutes the terms of each dot product).  Candidates: the caller's order, A's storage order,
  // B's storage order; keep the one that sends the fewest elements through the permute kernel.
  OperandPlan opa, opb;
  {
    std::vector<std::vector<int>> cands;
    std::vector<int> id(std::max(nctrct, 0));
    std::iota(id.begin(), id.end(), 0);
    cands.push_back(id);
    if (nctrct > 1 && per_block) {
      std::vector<int> sa = id, sb = id;
      const int32_t *ac = a_perm.data() + (a_rank - nctrct), *bc = b_perm.data();
      std::sort(sa.begin(), sa.end(), [&](int x, int y) { return ac[x] < ac[y]; });
      std::sort(sb.begin(), sb.end(), [&](int x, int y) { return bc[x] < bc[y]; });
      if (sa != id) cands.push_back(sa);
      if (sb != id && sb != sa) cands.push_back(sb);
    }
    uint64_t best_cost = ~0ull;
    std::vector<int32_t> best_a, best_b;
    for (const auto &sig : cands) {
      std::vector<int32_t> pa = a_perm, pb = b_perm;
      for (int i = 0; i < nctrct; ++i) { pa[a_rank - nctrct + i] = a_perm[a_rank - nctrct + sig[i]]; pb[i] = b_perm[sig[i]]; }
      OperandPlan ca = ClassifyOperand(a_rank, pa.data(), na, a_shape, a_used, a_rows, !is_ident(a_rank, pa.data()), per_block, allow_trans);
      OperandPlan cb = ClassifyOperand(b_rank, pb.data(), nb, b_shape, b_used, b_rows, !is_ident(b_rank, pb.data()), per_block, allow_trans,
                                       allow_trans && !(flags & QLB200_PLAN_NO_VIEW));
      const uint64_t cost = ca.permute_elems + cb.permute_elems;
      if (cost < best_cost) { best_cost = cost; best_a = pa; best_b = pb; opa = std::move(ca); opb = std::move(cb); }
    }
    a_perm = best_a; b_perm = best_b;
  }
  h->a_trans = opa.permute_elems > 0; h->b_trans = opb.permute_elems > 0;

  std::vector<uint64_t> a_new, b_new;
  std::string err;
  if (h->a_trans) {
    err = AddPermBlocks(h, a_rank, a_perm.data(), a_shape, a_off, opa.mode, a_used, 0, &a_new, &h->ws_a_elems, &h->permute_elems_a);
    if (!err.empty()) return err;
  }
  if (h->b_trans) {
    err = AddPermBlocks(h, b_rank, b_perm.data(), b_shape, b_off, opb.mode, b_used, 1, &b_new, &h->ws_b_elems, &h->permute_elems_b);
    if (!err.empty()) return err;
  }
  if (h->perm_tile_base.empty()) h->perm_tile_base.push_back(0);
  h->ws_off_a.assign(na, ~0ull); h->ws_off_b.assign(nb, ~0ull);
  for (uint64_t b = 0; b < na; ++b) if (h->a_trans && a_used[b] && opa.mode[b] == kBlkPermute) h->ws_off_a[b] = a_new[b];
  for (uint64_t b = 0; b < nb; ++b) if (h->b_trans && b_used[b] && opb.mode[b] == kBlkPermute) h->ws_off_b[b] = b_new[b];

  const uint64_t es = dtype == QLB200_C64 ? 16 : 8;
  const double fl = dtype == QLB200_C64 ? 8.0 : 2.0;
  // groups = runs of equal c_ord in the sorted task list
  for (size_t i = 0; i < st.size();) {
    size_t e = i;
    while (e < st.size() && st[e].c_ord == st[i].c_ord && st[e].c_off == st[i].c_off) ++e;
    GemmGroup g;
    std::memset(&g, 0, sizeof(g));
    g.c_off = st[i].c_off; g.m = st[i].m; g.n = st[i].n;
    g.task_begin = static_cast<uint32_t>(h->tasks.size());
    uint64_t ksum = 0;
    for (size_t t = i; t < e; ++t) {
      if (st[t].m != g.m || st[t].n != g.n) return "tasks of one output block disagree on m/n";
      GemmTask gt;
      const uint8_t ma = opa.mode[st[t].a_ord], mb = opb.mode[st[t].b_ord];
      gt.a_off = ma == kBlkPermute ? a_new[st[t].a_ord] : st[t].a_off;
      gt.b_off = mb == kBlkPermute ? b_new[st[t].b_ord] : st[t].b_off;
      gt.k = st[t].k; gt.sign = st[t].sign < 0 ? -1 : 1;
      gt.flags = uint16_t((ma != kBlkPermute ? kTaskASrc : 0) | (ma == kBlkTrans ? kTaskATrans : 0) |
                          (mb != kBlkPermute ? kTaskBSrc : 0) | (mb == kBlkTrans ? kTaskBTrans : 0));
      gt.b_rs = gt.b_run = g.n; gt.b_cs = 0; gt.pad_ = 0;          // row-major k x n (in place or permuted copy)
      if (mb == kBlkView) { gt.b_rs = opb.view[st[t].b_ord].rs; gt.b_cs = opb.view[st[t].b_ord].cs; gt.b_run = opb.view[st[t].b_ord].run; }
      h->tasks.push_back(gt);
      h->task_a_ord.push_back(st[t].a_ord);
      h->task_b_ord.push_back(st[t].b_ord);
      ksum += gt.k;
      h->flops += fl * double(g.m) * double(gt.k) * double(g.n);
      h->gemm_read_bytes += (uint64_t(g.m) * gt.k + uint64_t(gt.k) * g.n) * es;
    }
    h->gemm_write_bytes += uint64_t(g.m) * g.n * es;
    g.task_end = static_cast<uint32_t>(h->tasks.size());
    g.row_begin = 0; g.row_end = g.m;
    h->groups.push_back(g);
    h->group_ksum.push_back(ksum);
    i = e;
  }
  h->part_groups = h->groups;
  if (std::getenv("QLB200_DEBUG_TILES")) {
    unsigned cnt[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};   // tasks by operand mode: in place, 2-D transposed, strided view (B), permuted copy
    for (const GemmTask &t : h->tasks) {
      cnt[0][!(t.flags & kTaskASrc) ? 3 : (t.flags & kTaskATrans) ? 1 : 0]++;
      cnt[1][!(t.flags & kTaskBSrc) ? 3 : (t.flags & kTaskBTrans) ? 1 : (t.b_run < t.b_rs && t.b_cs != 0) ? 2 : 0]++;
    }
    std::fprintf(stderr, "[qlb200 plan] tasks %zu: A in place %u / transposed %u / permuted %u;  B in place %u / transposed %u / view %u / permuted %u\n",
                 h->tasks.size(), cnt[0][0], cnt[0][1], cnt[0][3], cnt[1][0], cnt[1][1], cnt[1][2], cnt[1][3]);
  }
  return BuildTiles(h);
}

namespace {

struct GroupClass { bool skinny; uint64_t kpad; };

GroupClass Classify(const PlanHost *h, const GemmGroup &g, int bk) {
  uint32_t kmax = 0;
  uint64_t kpad = 0;
  for (uint32_t t = g.task_begin; t < g.task_end; ++t) {
    kmax = std::max(kmax, h->tasks[t].k);
    kpad += (uint64_t(h->tasks[t].k) + bk - 1) / bk * bk;
  }
  const bool skinny = !(h->flags & QLB200_PLAN_NO_SKINNY) && g.n <= uint32_t(kSkinnyMaxN) && kmax <= uint32_t(kSkinnyMaxK);
  return {skinny, kpad};
}

}  // namespace

std::string BuildTiles(PlanHost *h) {
  h->tiles.clear(); h->items.clear(); h->seg.clear();
  h->n_split_ctrs = 0; h->n_part_slots = 0;
  int BM, BN, BK;
  const bool four_m = (h->flags & QLB200_PLAN_CPLX_4M) != 0;
  if (h->dtype == QLB200_C64) { BM = kWsBM; BN = four_m ? kWsBN : kWs3mBN; BK = four_m ? kWsBK : kWs3mBK; }
  else { BM = kWsRealBM; BN = kWsRealBN; BK = kWsRealBK; }
  h->part_slot_elems = uint64_t(BM) * BN;

  struct GInfo { uint32_t gi, tm, tn, stages; };
  std::vector<GInfo> dm;
  uint64_t total_stage_tiles = 0;
  {   // narrow-pair items: how many sub-chunks each may hold while every resident CTA (4 per SM) still gets ~2 items
    uint64_t chunks = 0;
    for (const GemmGroup &g : h->part_groups) {
      if (g.row_end <= g.row_begin || !Classify(h, g, 8).skinny) continue;
      const uint32_t per = uint32_t(kSkinnyElems) / g.n;
      chunks += (g.row_end - g.row_begin + per - 1) / per;
    }
    // items per resident CTA slot: 2 keeps the tail short and pays the descriptor chain + table build of an item (3-4 us, as
    // much as 1024 outputs take to stream) seldom; 8 was measured 25 % slower on 8-way shards, equal elsewhere (exp/r2_call34.sh)
    uint64_t per_slot = 2;
    if (const char *ov = std::getenv("QLB200_SKINNY_ITEMS_PER_SLOT")) per_slot = uint64_t(std::max(1, std::atoi(ov)));      // tuning aid
    const uint64_t want_items = uint64_t(std::max(1, h->num_sms)) * 4 * per_slot;
    h->skinny_sub = uint32_t(std::min<uint64_t>(kSkinnyMaxSub, std::max<uint64_t>(1, chunks / want_items)));
  }
  for (uint32_t gi = 0; gi < h->part_groups.size(); ++gi) {
    const GemmGroup &g = h->part_groups[gi];
    if (g.row_end <= g.row_begin) continue;
    const uint32_t rows = g.row_end - g.row_begin;
    if (Classify(h, g, 8).skinny) {
      const uint32_t per = uint32_t(kSkinnyElems) / g.n * h->skinny_sub;   // rows per work item (n <= kSkinnyMaxN)
      for (uint32_t r = 0; r < rows; r += per) h->items.push_back({gi, g.row_begin + r});
      continue;
    }
    const uint32_t tm = (rows + BM - 1) / BM, tn = (g.n + BN - 1) / BN;
    if (tm > 65535 || tn > 65535) return "output block too large for 16-bit tile coordinates";
    uint64_t stages = 0;
    for (uint32_t t = g.task_begin; t < g.task_end; ++t) stages += (uint64_t(h->tasks[t].k) + BK - 1) / BK;
    if (stages >= (1ull << 31)) return "k loop of one output block too long";
    dm.push_back({gi, tm, tn, uint32_t(stages)});
    total_stage_tiles += uint64_t(tm) * tn * stages;
  }
  // Split-K.  The dynamic scheduler balances the SMs well as long as no unit is long compared with a
  // CTA's share of the launch; when one output block has a very long k loop (multi-GPU row slabs,
  // small bond dimensions) its tiles are cut along k.  A cut is not free (the partial tile makes a
  // round trip through L2/HBM), so the cut length is chosen by simulating the greedy schedule for a
  // few candidates and keeping the shortest modelled makespan.
  const uint64_t slots = uint64_t(std::max(1, h->num_sms)) * 2;
  const int mt_full = BM / 8, nt_full = BN / 32;
  auto tile_weight = [&](const GInfo &d, uint32_t i, uint32_t j) {
    const GemmGroup &g = h->part_groups[d.gi];
    const uint32_t rows = std::min<uint32_t>(BM, g.row_end - g.row_begin - i * BM), cols = std::min<uint32_t>(BN, g.n - j * BN);
    const uint32_t mt = (rows + 7) / 8, nt = ((cols + 7) / 8 + 3) / 4;
    return double(mt * nt) / double(mt_full * nt_full);
  };
  constexpr double kSplitOverheadStages = 4.0;   // pipeline refill + partial-tile round trip, in stage units
  auto split_of = [](uint32_t stages, uint64_t chunk, uint32_t *len) {
    uint32_t nsplit = uint32_t(std::min<uint64_t>((stages + chunk - 1) / chunk, 64));
    if (nsplit < 1) nsplit = 1;
    *len = (stages + nsplit - 1) / nsplit;
    return (stages + *len - 1) / *len;
  };
  auto makespan = [&](uint64_t chunk) {
    std::vector<double> cost;
    for (const GInfo &d : dm) {
      uint32_t len;
      const uint32_t nsplit = split_of(d.stages, chunk, &len);
      for (uint32_t i = 0; i < d.tm; ++i)
        for (uint32_t j = 0; j < d.tn; ++j) {
          const double w = tile_weight(d, i, j);
          for (uint32_t sp = 0; sp < nsplit; ++sp) {
            const uint32_t n = std::min(d.stages, (sp + 1) * len) - sp * len;
            cost.push_back(n * w + (nsplit > 1 ? kSplitOverheadStages : 0.0));
          }
        }
    }
    std::sort(cost.begin(), cost.end(), std::greater<double>());
    std::vector<double> bins(slots, 0.0);
    auto cmp = [](double x, double y) { return x > y; };   // min-heap
    for (double cst : cost) {
      std::pop_heap(bins.begin(), bins.end(), cmp);
      bins.back() += cst;
      std::push_heap(bins.begin(), bins.end(), cmp);
    }
    return *std::max_element(bins.begin(), bins.end());
  };
  uint64_t chunk = ~0ull;
  if (!dm.empty() && !(h->flags & QLB200_PLAN_NO_SPLIT_K)) {
    constexpr uint64_t kMinChunk = 16;
    const uint64_t budget = std::max<uint64_t>(1, total_stage_tiles / slots);
    double best = makespan(~0ull);
    // a cut is only taken when it shortens the modelled makespan by 2 %; no schedule beats the perfectly divisible one, so an
    // uncut list already within 2 % of it needs no candidates (the usual case for whole problems: 6 simulations -> 1)
    double ideal = 0;
    for (const GInfo &d : dm)
      for (uint32_t i = 0; i < d.tm; ++i)
        for (uint32_t j = 0; j < d.tn; ++j) ideal += tile_weight(d, i, j) * d.stages;
    ideal /= double(slots);
    for (uint64_t div = 2; div <= 32 && best * 0.98 > ideal; div *= 2) {
      const uint64_t cand = std::max<uint64_t>(kMinChunk, budget / div);
      const double t = makespan(cand);
      if (t < best * 0.98) { best = t; chunk = cand; }
      if (cand == kMinChunk) break;
    }
  }
  if (!dm.empty() && (h->flags & QLB200_PLAN_STAGGER_OUTPUT) && !(h->flags & QLB200_PLAN_NO_SPLIT_K)) {
    uint32_t longest = 0;
    for (const GInfo &d : dm) longest = std::max(longest, d.stages);
    chunk = std::min<uint64_t>(chunk, std::max<uint64_t>(8, (longest + 3) / 4));
  }
  if (!dm.empty() && (h->flags & QLB200_PLAN_STREAM_K) && !(h->flags & QLB200_PLAN_NO_SPLIT_K)) {
    // Stream-K.  Tiles in block order (neighbours share operand panels in L2), their k loops laid end to end and weighted
    // by the share of MMA groups a tile issues; the line is cut into one equal-cost segment per resident CTA and CTA b runs
    // segment b (GemmParams::seg).  A tile that straddles a cut becomes 2+ units finished by the split-K fix-up; cuts that
    // would leave a piece shorter than kMinPiece stages snap to the tile boundary instead.
    constexpr uint32_t kMinPiece = 4;
    struct TileRef { uint32_t dm_idx, i, j; double w; };
    std::vector<TileRef> order;
    double W = 0;
    for (uint32_t x = 0; x < dm.size(); ++x)
      for (uint32_t i = 0; i < dm[x].tm; ++i)
        for (uint32_t j = 0; j < dm[x].tn; ++j) {
          const double w = tile_weight(dm[x], i, j);
          order.push_back({x, i, j, w});
          W += w * dm[x].stages;
        }
    // one segment per resident CTA, fewer only when the whole launch is tiny
    const uint32_t nseg = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>(slots, uint64_t(W / (2.0 * kMinPiece)))));
    double share = W / nseg, placed = 0;     // the share is re-derived from what is left whenever a segment closes
    h->seg.assign(1, 0u);
    double acc = 0;                 // cost already placed in the open segment
    uint32_t closed = 0;            // segments closed so far (the last one takes whatever is left)
    struct Piece { uint32_t end; bool closes; };
    std::vector<Piece> pieces;
    for (const TileRef &tr : order) {
      const GInfo &d = dm[tr.dm_idx];
      pieces.clear();
      uint32_t pos = 0;
      while (pos < d.stages) {
        const bool final_seg = closed + 1 >= nseg;
        const uint32_t rest = d.stages - pos;
        uint32_t take = rest;
        if (!final_seg) {
          const double room = share - acc;
          const double want = room > 0 ? std::floor(room / tr.w + 0.5) : 0.0;
          if (want < double(rest)) {
            take = uint32_t(want);
            if (take < kMinPiece) {
              if (pos == 0 && acc > 0) {              // the open segment is (almost) full: the tile starts the next one
                h->seg.push_back(uint32_t(h->tiles.size()));
                ++closed; acc = 0;
                share = (W - placed) / double(nseg - closed);
                continue;
              }
              take = std::min(kMinPiece, rest);
            }
            if (rest - take < kMinPiece) take = rest;  // no tiny tail piece either
          }
        }
        pos += take;
        acc += tr.w * take;
        placed += tr.w * take;
        const bool closes = !final_seg && acc + 0.5 * tr.w >= share;
        pieces.push_back({pos, closes});
        if (closes) { ++closed; acc = 0; share = (W - placed) / double(nseg - closed); }
      }
      const uint32_t nsplit = uint32_t(pieces.size());
      if (nsplit > 0xffffu) return "k loop cut into too many units";
      GemmTile t;
      std::memset(&t, 0, sizeof(t));
      t.group = d.gi; t.tm = uint16_t(tr.i); t.tn = uint16_t(tr.j); t.nsplit = uint16_t(nsplit);
      if (nsplit > 1) {
        t.ctr = h->n_split_ctrs++;
        t.part_base = static_cast<uint32_t>(h->n_part_slots);
        h->n_part_slots += nsplit;
      }
      uint32_t begin = 0;
      for (uint32_t sp = 0; sp < nsplit; ++sp) {
        t.split = uint16_t(sp);
        t.s_begin = begin; t.s_end = pieces[sp].end;
        begin = pieces[sp].end;
        h->tiles.push_back(t);
        if (pieces[sp].closes) h->seg.push_back(uint32_t(h->tiles.size()));
      }
    }
    if (h->seg.back() != h->tiles.size()) h->seg.push_back(uint32_t(h->tiles.size()));
    if (h->n_part_slots >= (1ull << 32)) return "too many split-K slots";
    if (h->tiles.size() >= (1ull << 32) || h->items.size() >= (1ull << 32)) return "too many tiles";
    return "";
  }
  if (const char *ov = std::getenv("QLB200_SPLIT_CHUNK")) {      // tuning aid: force the split-K cut length (stages)
    const long long v = std::atoll(ov);
    if (v > 0) chunk = uint64_t(v);
  }
  if (std::getenv("QLB200_DEBUG_TILES") && !h->items.empty())
    std::fprintf(stderr, "[qlb200 tiles] narrow-pair items %zu, sub-chunks per item %u\n", h->items.size(), h->skinny_sub);
  if (std::getenv("QLB200_DEBUG_TILES") && !dm.empty()) {
    double wsum = 0;
    for (const GInfo &d : dm)
      for (uint32_t i = 0; i < d.tm; ++i)
        for (uint32_t j = 0; j < d.tn; ++j) wsum += tile_weight(d, i, j) * d.stages;
    std::fprintf(stderr, "[qlb200 tiles] groups %zu stage-tiles %llu weighted %.0f ideal/slot %.1f chunk %lld makespan(nosplit) %.1f makespan(chosen) %.1f\n",
                 dm.size(), (unsigned long long) total_stage_tiles, wsum, wsum / double(slots), chunk == ~0ull ? -1ll : (long long) chunk,
                 makespan(~0ull), makespan(chunk));
  }
  for (const GInfo &d : dm) {
    uint32_t len;
    const uint32_t nsplit = split_of(d.stages, chunk, &len);
    for (uint32_t i = 0; i < d.tm; ++i)
      for (uint32_t j = 0; j < d.tn; ++j) {
        GemmTile t;
        std::memset(&t, 0, sizeof(t));
        t.group = d.gi; t.tm = uint16_t(i); t.tn = uint16_t(j); t.nsplit = uint16_t(nsplit);
        if (nsplit > 1) {
          t.ctr = h->n_split_ctrs++;
          t.part_base = static_cast<uint32_t>(h->n_part_slots);
          h->n_part_slots += nsplit;
        }
        for (uint32_t sp = 0; sp < nsplit; ++sp) {
          t.split = uint16_t(sp);
          t.s_begin = sp * len; t.s_end = std::min(d.stages, (sp + 1) * len);
          h->tiles.push_back(t);
        }
      }
  }
  if (h->n_part_slots >= (1ull << 32)) return "too many split-K slots";
  // costliest units first: persistent CTAs then finish with the cheap ones (LPT).  Cost = k-loop length x the
  // fraction of the tile's MMAs that are issued (ragged edge tiles skip the groups outside the block).
  {
    std::vector<uint32_t> gi_to_dm(h->part_groups.size(), 0);
    for (uint32_t x = 0; x < dm.size(); ++x) gi_to_dm[dm[x].gi] = x;
    std::vector<std::pair<double, uint32_t>> key(h->tiles.size());
    for (uint32_t x = 0; x < h->tiles.size(); ++x) {
      const GemmTile &t = h->tiles[x];
      key[x] = {-double(t.s_end - t.s_begin) * tile_weight(dm[gi_to_dm[t.group]], t.tm, t.tn), x};
    }
    std::stable_sort(key.begin(), key.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    std::vector<GemmTile> sorted(h->tiles.size());
    for (uint32_t x = 0; x < key.size(); ++x) sorted[x] = h->tiles[key[x].second];
    h->tiles.swap(sorted);
  }
  if (h->tiles.size() >= (1ull << 32) || h->items.size() >= (1ull << 32)) return "too many tiles";
  return "";
}

void PartitionRows(PlanHost *h, int world, int rank) {
  // One line of output rows (groups in C order), each row weighted by its flops; rank r owns the
  // r-th equal-cost segment.  Cuts are snapped to 64-row boundaries inside a group so that MMA
  // tiles are not split unevenly.  Every rank computes the same cuts, no communication needed.
  const size_t ng = h->groups.size();
  std::vector<double> row_cost(ng);
  double total = 0;
  for (size_t i = 0; i < ng; ++i) {
    row_cost[i] = double(h->groups[i].n) * double(h->group_ksum[i]) + 1e-9;
    total += row_cost[i] * h->groups[i].m;
  }
  auto cut_pos = [&](double target, size_t *gi, uint32_t *row) {
    double acc = 0;
    for (size_t i = 0; i < ng; ++i) {
      const double gc = row_cost[i] * h->groups[i].m;
      if (acc + gc > target) {
        uint32_t r = static_cast<uint32_t>((target - acc) / row_cost[i]);
        r = (r + 32) / 64 * 64;
        if (r > h->groups[i].m) r = h->groups[i].m;
        *gi = i; *row = r;
        return;
      }
      acc += gc;
    }
    *gi = ng; *row = 0;
  };
  size_t g0 = 0, g1 = ng; uint32_t r0 = 0, r1 = 0;
  if (rank > 0) cut_pos(total * rank / world, &g0, &r0);
  if (rank < world - 1) cut_pos(total * (rank + 1) / world, &g1, &r1);
  h->part_groups = h->groups;
  for (size_t i = 0; i < ng; ++i) {
    GemmGroup &g = h->part_groups[i];
    uint32_t lo = 0, hi = g.m;
    if (i < g0) hi = 0;
    else if (i == g0) lo = r0;
    if (i > g1) hi = 0;
    else if (i == g1) hi = std::min(hi, r1);
    if (hi < lo) hi = lo;
    g.row_begin = lo; g.row_end = hi;
  }
  FilterPermBlocks(h);
}

void FilterPermBlocks(PlanHost *h) {
  // a rank permutes only the operand blocks its share of the output rows reads (whole blocks: B blocks are shared by all
  // rows of an output block, A blocks are cut along m only by a row slab, which does not pay for a second descriptor)
  if (h->perm_blks_all.empty()) return;
  std::vector<char> a_need, b_need;
  for (size_t gi = 0; gi < h->part_groups.size(); ++gi) {
    const GemmGroup &g = h->part_groups[gi];
    if (g.row_end <= g.row_begin) continue;
    for (uint32_t t = g.task_begin; t < g.task_end; ++t) {
      if (h->task_a_ord[t] >= a_need.size()) a_need.resize(h->task_a_ord[t] + 1, 0);
      if (h->task_b_ord[t] >= b_need.size()) b_need.resize(h->task_b_ord[t] + 1, 0);
      a_need[h->task_a_ord[t]] = 1; b_need[h->task_b_ord[t]] = 1;
    }
  }
  h->perm_blks.clear();
  h->perm_tile_base.assign(1, 0u);
  h->permute_elems_a = h->permute_elems_b = 0;
  for (size_t i = 0; i < h->perm_blks_all.size(); ++i) {
    const bool is_b = (h->perm_owner_all[i] >> 63) != 0;
    const uint64_t ord = h->perm_owner_all[i] & ~(1ull << 63);
    const std::vector<char> &need = is_b ? b_need : a_need;
    if (ord >= need.size() || !need[ord]) continue;
    h->perm_blks.push_back(h->perm_blks_all[i]);
    h->perm_tile_base.push_back(h->perm_tile_base.back() + h->perm_ntiles_all[i]);
    (is_b ? h->permute_elems_b : h->permute_elems_a) += h->perm_size_all[i];
  }
}

}  // namespace qlb200
