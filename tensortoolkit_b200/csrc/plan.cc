// plan.cc -- host-side construction of the device descriptor tables (permute blocks, GEMM groups,
// tile lists, multi-GPU row partition).  No reference counterpart: the reference walks tasks one
// by one (global_operations.h:919-982); here the whole contraction is flattened into tables once.
#include "plan.h"

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
                    uint32_t src_sel, float scale, uint64_t *ntiles_out) {
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

static std::string AddPermBlocks(PlanHost *h, int rank, const int32_t *perm, const uint32_t *shape,
                                 const uint64_t *off, const std::vector<char> &used, uint32_t src_sel,
                                 std::vector<uint64_t> *new_off, uint64_t *ws_elems, uint64_t *moved) {
  const uint64_t n = used.size();
  new_off->assign(n, 0);
  uint64_t ws = 0;
  for (uint64_t b = 0; b < n; ++b) {
    if (!used[b]) continue;
    uint64_t sz = 1;
    for (int i = 0; i < rank; ++i) sz *= shape[b * rank + i];
    if (sz >= (1ull << 32)) return "block with 2^32 or more elements";
    (*new_off)[b] = ws;
    uint64_t nt = 0;
    PermBlk d = MakePermBlk(rank, shape + b * rank, perm, off[b], ws, src_sel, 1.0f, &nt);
    const uint64_t base = h->perm_tile_base.empty() ? 0 : h->perm_tile_base.back();
    if (h->perm_tile_base.empty()) h->perm_tile_base.push_back(0);
    if (base + nt >= (1ull << 32)) return "too many permute tiles";
    h->perm_blks.push_back(d);
    h->perm_tile_base.push_back(static_cast<uint32_t>(h->perm_tile_base.back() + nt));
    ws += (sz + 1) & ~1ull;        // keep every permuted block 16-byte aligned for doubles
    *moved += sz;
  }
  *ws_elems = ws;
  return "";
}

std::string BuildPlanHost(int dtype, uint32_t flags, bool a_trans, int a_rank, const int32_t *a_perm,
                          uint64_t na, const uint32_t *a_shape, const uint64_t *a_off, uint64_t a_elems,
                          bool b_trans, int b_rank, const int32_t *b_perm, uint64_t nb, const uint32_t *b_shape,
                          const uint64_t *b_off, uint64_t b_elems, const std::vector<qlb200_task> &st,
                          uint64_t c_elems, PlanHost *h) {
  h->dtype = dtype; h->flags = flags;
  h->a_elems = a_elems; h->b_elems = b_elems; h->c_elems = c_elems;
  h->a_trans = a_trans; h->b_trans = b_trans;
  std::vector<char> a_used(na, 0), b_used(nb, 0);
  for (const auto &t : st) {
    if (t.a_ord >= na || t.b_ord >= nb) return "task references a block ordinal out of range";
    a_used[t.a_ord] = 1; b_used[t.b_ord] = 1;
  }
  std::vector<uint64_t> a_new, b_new;
  std::string err;
  if (a_trans) {
    err = AddPermBlocks(h, a_rank, a_perm, a_shape, a_off, a_used, 0, &a_new, &h->ws_a_elems, &h->permute_elems_a);
    if (!err.empty()) return err;
  }
  if (b_trans) {
    err = AddPermBlocks(h, b_rank, b_perm, b_shape, b_off, b_used, 1, &b_new, &h->ws_b_elems, &h->permute_elems_b);
    if (!err.empty()) return err;
  }
  if (h->perm_tile_base.empty()) h->perm_tile_base.push_back(0);

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
      gt.a_off = a_trans ? a_new[st[t].a_ord] : st[t].a_off;
      gt.b_off = b_trans ? b_new[st[t].b_ord] : st[t].b_off;
      gt.k = st[t].k; gt.sign = st[t].sign < 0 ? -1 : 1;
      h->tasks.push_back(gt);
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
  h->tiles.clear(); h->items.clear();
  int BM = kRealBM, BN = kRealBN;
  if (h->dtype == QLB200_C64) {
    const bool legacy = (h->flags & QLB200_PLAN_LEGACY_GEMM) != 0;
    BM = legacy ? kCplxBM : kWsBM; BN = legacy ? kCplxBN : kWsBN;
  }
  std::vector<uint32_t> order(h->part_groups.size());
  std::iota(order.begin(), order.end(), 0u);
  // heaviest k-loops first: persistent CTAs then finish with the short tiles (LPT)
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return h->group_ksum[x] > h->group_ksum[y]; });
  for (uint32_t gi : order) {
    const GemmGroup &g = h->part_groups[gi];
    if (g.row_end <= g.row_begin) continue;
    const uint32_t rows = g.row_end - g.row_begin;
    if (Classify(h, g, 8).skinny) {
      for (uint32_t r = 0; r < rows; r += kSkinnyRows) h->items.push_back({gi, g.row_begin + r});
    } else {
      const uint32_t tm = (rows + BM - 1) / BM, tn = (g.n + BN - 1) / BN;
      if (tm > 65535 || tn > 65535) return "output block too large for 16-bit tile coordinates";
      for (uint32_t i = 0; i < tm; ++i)
        for (uint32_t j = 0; j < tn; ++j) h->tiles.push_back({gi, uint16_t(i), uint16_t(j)});
    }
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
}

}  // namespace qlb200
