// qlten_b200/sharding.h -- the host side of the multi-GPU path for TensorToolkit's own tensor types.
//
// The reference distributes the DMRG mat-vec over MPI ranks by restricting ONE free index of the first operand to one QN
// sector per work unit (qlten::dmrg::Contract1Sector, tensor_manipulation/dmrg/contract_1sector.h:181-228).  The B200 path
// cuts that index by ROWS (include/qlb200.h, "row-slab partitioner"): a rank's operand is the row slab below.
//
//   qlten::b200::SectorFlops(a, b, axes_set, split_axis, cost)   flops of Contract(a, b, axes_set) per sector of a's index
//                                                                `split_axis` (added to cost; sum over the chain's steps)
//   qlten::b200::CutRowLine(cost, degeneracies, world, snap)     rows [lo, hi) of every sector owned by every rank
//   qlten::b200::RowSlab(t, axis, ranges)                        the tensor restricted to those rows: same QN sectors with
//                                                                reduced degeneracies (empty ones dropped), blocks sliced
//
// Only public reference API is used (as in contract.h).
#ifndef QLTEN_B200_SHARDING_H
#define QLTEN_B200_SHARDING_H

#include <cstring>
#include <utility>
#include <vector>

#include "qlten_b200/contract.h"

namespace qlten {
namespace b200 {

using RowRanges = std::vector<std::pair<uint32_t, uint32_t>>;   // per sector of the split index: rows [first, second)

template<typename TenElemT, typename QNT>
void SectorFlops(const QLTensor<TenElemT, QNT> &a, const QLTensor<TenElemT, QNT> &b, const std::vector<std::vector<size_t>> &axes_set,
                 size_t split_axis, std::vector<double> &cost) {
  if (split_axis >= a.Rank()) { throw std::invalid_argument("b200::SectorFlops: bad split axis"); }
  cost.resize(a.GetIndex(split_axis).GetQNSctNum(), 0.0);
  detail::ShellHolder sa, sb;
  detail::FillShell(a, sa);
  detail::FillShell(b, sb);
  std::vector<int32_t> aa(axes_set[0].begin(), axes_set[0].end()), ba(axes_set[1].begin(), axes_set[1].end());
  detail::MatchGuard match;
  detail::Check(qlb200_match_create(&sa.shell, &sb.shell, static_cast<int32_t>(aa.size()), aa.data(), ba.data(), &match.m), "match_create");
  detail::Check(qlb200_shard_sector_flops(match.m, static_cast<int32_t>(split_axis), detail::DTypeOf<TenElemT>::value, cost.data()),
                "shard_sector_flops");
}

/// result[rank][sector] = rows [lo, hi); every rank computes the same cuts from the same numbers
inline std::vector<RowRanges> CutRowLine(const std::vector<double> &cost, const std::vector<uint32_t> &degs, int world, int snap = 8) {
  const uint32_t nsct = static_cast<uint32_t>(degs.size());
  std::vector<qlb200_piece> line(nsct);
  for (uint32_t s = 0; s < nsct; ++s) {
    line[s].sector = s; line[s].lo = 0; line[s].hi = degs[s]; line[s].pad_ = 0;
    line[s].weight = degs[s] ? cost[s] / degs[s] : 0.0;
  }
  std::vector<uint32_t> flat(static_cast<size_t>(world) * nsct * 2);
  detail::Check(qlb200_shard_cut_line(line.data(), nsct, degs.data(), nsct, world, snap, flat.data()), "shard_cut_line");
  std::vector<RowRanges> out(world, RowRanges(nsct));
  for (int r = 0; r < world; ++r) {
    for (uint32_t s = 0; s < nsct; ++s) {
      out[r][s] = {flat[(static_cast<size_t>(r) * nsct + s) * 2], flat[(static_cast<size_t>(r) * nsct + s) * 2 + 1]};
    }
  }
  return out;
}

template<typename TenElemT, typename QNT>
QLTensor<TenElemT, QNT> RowSlab(const QLTensor<TenElemT, QNT> &t, size_t axis, const RowRanges &ranges) {
  if (axis >= t.Rank() || ranges.size() != t.GetIndex(axis).GetQNSctNum()) {
    throw std::invalid_argument("b200::RowSlab: bad axis / one range per sector of the index expected");
  }
  detail::ShellHolder h;
  detail::FillShell(t, h);
  std::vector<uint32_t> flat;
  for (const auto &r : ranges) { flat.push_back(r.first); flat.push_back(r.second); }
  qlb200_slab_info info;
  detail::Check(qlb200_shard_restrict(&h.shell, static_cast<int32_t>(axis), flat.data(), &info, nullptr, nullptr, nullptr, nullptr, nullptr,
                                      nullptr, nullptr), "shard_restrict");
  const size_t rank = t.Rank();
  std::vector<uint32_t> kept(info.nsct_kept + 1), ndeg(info.nsct_kept + 1), kblk(info.nblk_kept + 1), ncoor((info.nblk_kept + 1) * rank);
  std::vector<uint64_t> src(info.ncopy + 1), dst(info.ncopy + 1), len(info.ncopy + 1);
  detail::Check(qlb200_shard_restrict(&h.shell, static_cast<int32_t>(axis), flat.data(), &info, kept.data(), ndeg.data(), kblk.data(),
                                      ncoor.data(), src.data(), dst.data(), len.data()), "shard_restrict");
  const Index<QNT> &old = t.GetIndex(axis);
  QNSectorVec<QNT> scts;
  for (uint32_t i = 0; i < info.nsct_kept; ++i) { scts.push_back(QNSector<QNT>(old.GetQNSct(kept[i]).GetQn(), ndeg[i])); }
  IndexVec<QNT> indexes = t.GetIndexes();
  indexes[axis] = Index<QNT>(scts, old.GetDir());
  QLTensor<TenElemT, QNT> out(indexes);
  if (info.nblk_kept == 0) { return out; }
  std::vector<size_t> nsct(rank);
  for (size_t i = 0; i < rank; ++i) { nsct[i] = indexes[i].GetQNSctNum(); }
  std::vector<size_t> idxs(info.nblk_kept);
  std::vector<CoorsT> coors(info.nblk_kept, CoorsT(rank));
  for (uint64_t b = 0; b < info.nblk_kept; ++b) {
    size_t idx = 0;
    for (size_t i = 0; i < rank; ++i) {
      coors[b][i] = ncoor[b * rank + i];
      idx = idx * nsct[i] + coors[b][i];                    // row-major block index over the slab's sector counts
    }
    idxs[b] = idx;
  }
  auto &bsdt = out.GetBlkSparDataTen();
  bsdt.DataBlksInsert(idxs, coors, true);                   // sets offsets + raw_data_size_, allocates (uninitialised)
  const TenElemT *from = t.GetBlkSparDataTen().GetActualRawDataPtr();
  TenElemT *to = bsdt.GetActualRawDataPtr();
  for (uint64_t k = 0; k < info.ncopy; ++k) { std::memcpy(to + dst[k], from + src[k], len[k] * sizeof(TenElemT)); }
  return out;
}

}  // namespace b200
}  // namespace qlten
#endif  // QLTEN_B200_SHARDING_H
