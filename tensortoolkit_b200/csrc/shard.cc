// shard.cc -- row-slab partitioner of a contraction chain over the GPUs of one NVSwitch domain (host only, no CUDA calls).
//
// The reference distributes the DMRG mat-vec over MPI ranks by restricting one FREE index of the first operand to a single
// QN sector per work unit (dmrg::Contract1Sector, tensor_manipulation/dmrg/contract_1sector.h:181-228) and summing the
// partial results.  Here the work unit is a RANGE OF ROWS of one sector of that index: all rows of the split index form one
// line (sector-major), weighted by the flops they cause in every step of the chain; rank r owns the r-th equal-weight
// segment.  Every rank computes the same cuts from the same numbers, so nothing has to be communicated.
//   qlb200_shard_sector_flops  (capi.cu) flops of one contraction attributed to the sectors of the split index
//   qlb200_shard_cut_line      the equal-weight cut of a line of weighted pieces (cuts snapped to multiples of `snap` rows)
//   qlb200_shard_reweigh       feedback step: pieces re-weighted by (measured time / modelled weight)^damp of their rank
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "qlb200.h"

extern "C" {

int qlb200_shard_cut_line(const qlb200_piece *pieces, uint64_t npieces, const uint32_t *degs, uint32_t nsct, int32_t world,
                          int32_t snap, uint32_t *ranges_out) {
  if ((!pieces && npieces) || !degs || !ranges_out || world < 1 || snap < 1) return QLB200_ERR_ARG;
  double total = 0.0;
  for (uint64_t i = 0; i < npieces; ++i) {
    if (pieces[i].sector >= nsct || pieces[i].hi < pieces[i].lo || pieces[i].hi > degs[pieces[i].sector]) return QLB200_ERR_ARG;
    total += double(pieces[i].hi - pieces[i].lo) * pieces[i].weight;
  }
  struct Cut { uint32_t sector, row; };
  auto locate = [&](double target) {
    double acc = 0.0;
    for (uint64_t i = 0; i < npieces; ++i) {
      const qlb200_piece &p = pieces[i];
      const double c = double(p.hi - p.lo) * p.weight;
      if (acc + c > target && c > 0) {
        double r = double(p.lo) + (target - acc) / p.weight;
        r = std::nearbyint(r / double(snap)) * double(snap);          // round half to even, like the Python restatement
        const double hi = double(degs[p.sector]);
        return Cut{p.sector, uint32_t(std::max(0.0, std::min(hi, r)))};
      }
      acc += c;
    }
    return Cut{nsct, 0u};
  };
  std::vector<Cut> cuts;
  cuts.push_back(Cut{0u, 0u});
  for (int32_t r = 1; r < world; ++r) cuts.push_back(locate(total * double(r) / double(world)));
  cuts.push_back(Cut{nsct, 0u});
  for (size_t i = 1; i < cuts.size(); ++i)                            // snapping must not make the cuts run backwards
    if (cuts[i].sector < cuts[i - 1].sector || (cuts[i].sector == cuts[i - 1].sector && cuts[i].row < cuts[i - 1].row)) cuts[i] = cuts[i - 1];
  for (int32_t r = 0; r < world; ++r) {
    const Cut c0 = cuts[r], c1 = cuts[r + 1];
    for (uint32_t s = 0; s < nsct; ++s) {
      uint32_t lo = 0, hi = degs[s];
      if (s < c0.sector || s > c1.sector) {
        lo = hi = 0;
      } else {
        if (s == c0.sector) lo = c0.row;
        if (s == c1.sector) hi = c1.row;
      }
      uint32_t *o = ranges_out + (uint64_t(r) * nsct + s) * 2;
      o[0] = lo; o[1] = std::max(lo, hi);
    }
  }
  return QLB200_OK;
}

uint64_t qlb200_shard_reweigh(const qlb200_piece *pieces, uint64_t npieces, const uint32_t *ranges, uint32_t nsct, int32_t world,
                              const double *times, double damp, uint64_t cap, qlb200_piece *out) {
  if ((!pieces && npieces) || !ranges || !times || world < 1) return 0;
  auto overlap = [&](const qlb200_piece &p, int32_t r, uint32_t *a, uint32_t *b) {
    const uint32_t *rg = ranges + (uint64_t(r) * nsct + p.sector) * 2;
    *a = std::max(p.lo, rg[0]); *b = std::min(p.hi, rg[1]);
    return *b > *a;
  };
  std::vector<double> model(world, 0.0);
  for (int32_t r = 0; r < world; ++r)
    for (uint64_t i = 0; i < npieces; ++i) {
      uint32_t a, b;
      if (pieces[i].sector < nsct && overlap(pieces[i], r, &a, &b)) model[r] += double(b - a) * pieces[i].weight;
    }
  double sum_t = 0.0, sum_m = 0.0;
  int32_t n_t = 0, n_m = 0;
  for (int32_t r = 0; r < world; ++r)
    if (model[r] > 0) { sum_t += times[r]; ++n_t; sum_m += model[r]; ++n_m; }
  double mean_t = n_t ? sum_t / n_t : 1.0, mean_m = n_m ? sum_m / n_m : 1.0;
  if (mean_t == 0.0) mean_t = 1.0;
  if (mean_m == 0.0) mean_m = 1.0;
  std::vector<qlb200_piece> res;
  for (uint64_t i = 0; i < npieces; ++i) {
    if (pieces[i].sector >= nsct) continue;
    for (int32_t r = 0; r < world; ++r) {
      uint32_t a, b;
      if (!overlap(pieces[i], r, &a, &b)) continue;
      const double f = model[r] > 0 ? (times[r] / mean_t) / (model[r] / mean_m) : 1.0;
      qlb200_piece q = pieces[i];
      q.lo = a; q.hi = b; q.weight = pieces[i].weight * std::pow(f, damp);
      res.push_back(q);
    }
  }
  std::stable_sort(res.begin(), res.end(), [](const qlb200_piece &x, const qlb200_piece &y) {
    return x.sector != y.sector ? x.sector < y.sector : x.lo < y.lo;
  });
  for (uint64_t i = 0; i < res.size() && i < cap && out; ++i) out[i] = res[i];
  return res.size();
}

}  // extern "C"

// Row slab of a tensor: keep rows [lo, hi) of every sector of index `axis`; sectors with hi <= lo disappear and their blocks
// with them.  The relabelling of the kept sectors is monotone, so the kept blocks stay in ascending blk_idx order and the slab's
// raw buffer is the kept blocks' slices packed in that order.  The slice of a block is `outer` runs of (hi - lo) * inner
// contiguous elements (outer / inner = the extents in front of / behind `axis`): the copy list below builds the slab's buffer
// from the full one on the host or, through qlb200_cplan_create, on the device.
extern "C" int qlb200_shard_restrict(const qlb200_shell *t, int32_t axis, const uint32_t *ranges, qlb200_slab_info *info,
                                     uint32_t *kept_sectors, uint32_t *new_deg, uint32_t *kept_blocks, uint32_t *new_coors,
                                     uint64_t *copy_src, uint64_t *copy_dst, uint64_t *copy_len) {
  if (!t || !ranges || !info || axis < 0 || axis >= t->rank || !t->nsct || !t->deg || (t->nblk && !t->blk_coors)) return QLB200_ERR_ARG;
  const int rank = t->rank;
  std::vector<uint32_t> base(rank + 1, 0);
  for (int i = 0; i < rank; ++i) base[i + 1] = base[i] + t->nsct[i];
  const uint32_t nsct = t->nsct[axis];
  std::vector<int64_t> new_pos(nsct, -1);
  uint32_t nkept = 0;
  for (uint32_t s = 0; s < nsct; ++s) {
    const uint32_t lo = ranges[2 * s], hi = ranges[2 * s + 1];
    if (hi > t->deg[base[axis] + s]) return QLB200_ERR_ARG;
    if (hi > lo) {
      if (kept_sectors) kept_sectors[nkept] = s;
      if (new_deg) new_deg[nkept] = hi - lo;
      new_pos[s] = nkept++;
    }
  }
  uint64_t nblk = 0, elems = 0, ncopy = 0, src_off = 0;
  for (uint64_t b = 0; b < t->nblk; ++b) {
    const uint32_t *c = t->blk_coors + b * rank;
    uint64_t outer = 1, inner = 1, size = 1;
    for (int i = 0; i < rank; ++i) {
      if (c[i] >= t->nsct[i]) return QLB200_ERR_ARG;
      const uint64_t d = t->deg[base[i] + c[i]];
      size *= d;
      if (i < axis) outer *= d;
      if (i > axis) inner *= d;
    }
    const uint32_t s = c[axis];
    if (new_pos[s] >= 0) {
      const uint64_t lo = ranges[2 * s], hi = ranges[2 * s + 1], rows = t->deg[base[axis] + s];
      if (kept_blocks) kept_blocks[nblk] = uint32_t(b);
      if (new_coors) {
        for (int i = 0; i < rank; ++i) new_coors[nblk * rank + i] = c[i];
        new_coors[nblk * rank + axis] = uint32_t(new_pos[s]);
      }
      const uint64_t run = (hi - lo) * inner;
      if (hi - lo == rows || outer == 1) {                 // one contiguous piece
        if (copy_src) { copy_src[ncopy] = src_off + lo * inner; copy_dst[ncopy] = elems; copy_len[ncopy] = run * outer; }
        ++ncopy;
      } else {
        for (uint64_t o = 0; o < outer; ++o) {
          if (copy_src) { copy_src[ncopy] = src_off + (o * rows + lo) * inner; copy_dst[ncopy] = elems + o * run; copy_len[ncopy] = run; }
          ++ncopy;
        }
      }
      elems += run * outer;
      ++nblk;
    }
    src_off += size;
  }
  info->nsct_kept = nkept; info->nblk_kept = nblk; info->elems = elems; info->ncopy = ncopy;
  return QLB200_OK;
}
