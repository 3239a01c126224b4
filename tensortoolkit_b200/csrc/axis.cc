// axis.cc -- host side of the matrix-free axis operations: block pairing and output topology of
//   qlten::dmrg::ApplyRank2ToAxisPreserveOrder      (tensor_manipulation/dmrg/axis_ops.h:2889-2992)
//   qlten::dmrg::ApplyTwoRank2ToAxesPreserveOrder   (:2994-3125)
// i.e. out[.., j1, .., j2, ..] = sum_{i1, i2} in[.., i1, .., i2, ..] * op1[i1, j1] * op2[i2, j2]  with the axis ORDER kept.
// Behavioural contract: the output indexes (:905-914, :1777-1789), the output block set (GenerateRank2OutputBlockTopology
// :984-1027, GenerateTwoRank2OutputBlockTopology :1791-1841: every (input block, matching op block[s]) reaches one output
// block; blocks in ascending blk_idx order, DataBlksInsert layout) and the accumulation of every such triple into its
// output block (:2944-2990, :3056-3123).  The reference walks the triples and calls one GEMM each; here they are flattened
// into a term table per OUTPUT block, consumed by one kernel launch (axis.cu).
#include "matcher.h"

#include <algorithm>
#include <cstring>
#include <map>

namespace qlb200 {

std::string BuildAxisMatch(const qlb200_shell *sin, int nops, const qlb200_shell *sop1, int axis1, const qlb200_shell *sop2, int axis2,
                           AxisMatch *out) {
  AxisMatch &m = *out;
  if (nops != 1 && nops != 2) return "one or two rank-2 operators";
  std::string err = m.in.Load(sin);
  if (!err.empty()) return "input: " + err;
  err = m.op[0].Load(sop1);
  if (!err.empty()) return "op1: " + err;
  if (nops == 2) {
    err = m.op[1].Load(sop2);
    if (!err.empty()) return "op2: " + err;
  }
  m.nops = nops; m.axis[0] = axis1; m.axis[1] = nops == 2 ? axis2 : -1;
  const Shell &in = m.in;
  if (in.rank < 1) return "input must have at least one index";
  if (in.fermionic()) return "axis operations are bosonic-only (the reference static_asserts it)";
  if (nops == 2 && axis1 == axis2) return "axes must be distinct";
  for (int o = 0; o < nops; ++o) {
    const Shell &op = m.op[o];
    const int ax = m.axis[o];
    if (ax < 0 || ax >= in.rank) return "target axis out of range";
    if (op.rank != 2) return "rank2_op must have rank 2";
    if (op.fermionic()) return "axis operations are bosonic-only";
    if (op.nsct[0] != in.nsct[ax]) return "rank2_op input index must be the inverse of the tensor axis (sector count)";
    for (uint32_t s = 0; s < op.nsct[0]; ++s)
      if (op.deg[op.sct_base[0] + s] != in.deg[in.sct_base[ax] + s]) return "rank2_op input index must be the inverse of the tensor axis (degeneracy)";
  }
  // output index set: the input's, the target axes replaced by the operators' output indexes
  const int r = in.rank;
  m.out_nsct.assign(in.nsct.begin(), in.nsct.end());
  for (int o = 0; o < nops; ++o) m.out_nsct[m.axis[o]] = m.op[o].nsct[1];
  auto out_deg = [&](int ax, uint32_t s) -> uint32_t {
    for (int o = 0; o < nops; ++o)
      if (ax == m.axis[o]) return m.op[o].deg[m.op[o].sct_base[1] + s];
    return in.deg[in.sct_base[ax] + s];
  };
  // op blocks by input sector
  std::vector<std::vector<uint32_t>> by_sector[2];
  for (int o = 0; o < nops; ++o) {
    by_sector[o].assign(m.op[o].nsct[0], {});
    for (uint64_t b = 0; b < m.op[o].nblk; ++b) by_sector[o][m.op[o].coors[b * 2]].push_back(uint32_t(b));
  }
  // walk the (input block, op1 block[, op2 block]) triples in the reference's order; group them by output block
  std::map<uint64_t, std::vector<AxisTerm>> groups;
  std::map<uint64_t, std::vector<uint32_t>> coors_of;
  static const std::vector<uint32_t> kNone(1, 0xffffffffu);
  uint32_t c[QLB200_MAX_RANK];
  for (uint64_t ib = 0; ib < in.nblk; ++ib) {
    const std::vector<uint32_t> &l1 = by_sector[0][in.coors[ib * r + m.axis[0]]];
    const std::vector<uint32_t> &l2 = nops == 2 ? by_sector[1][in.coors[ib * r + m.axis[1]]] : kNone;
    for (uint32_t b1 : l1)
      for (uint32_t b2 : l2) {
        for (int i = 0; i < r; ++i) c[i] = in.coors[ib * r + i];
        c[m.axis[0]] = m.op[0].coors[b1 * 2 + 1];
        if (nops == 2) c[m.axis[1]] = m.op[1].coors[b2 * 2 + 1];
        uint64_t idx = 0;
        for (int i = 0; i < r; ++i) idx = idx * m.out_nsct[i] + c[i];
        auto &g = groups[idx];
        if (g.empty()) coors_of[idx].assign(c, c + r);
        g.push_back({uint32_t(ib), b1, nops == 2 ? b2 : 0u});
      }
  }
  uint64_t off = 0;
  for (auto &kv : groups) {
    CBlock b;
    std::memset(&b, 0, sizeof(b));
    b.blk_idx = kv.first; b.offset = off; b.size = 1;
    const std::vector<uint32_t> &cc = coors_of[kv.first];
    for (int i = 0; i < r; ++i) { b.coors[i] = cc[i]; b.shape[i] = out_deg(i, cc[i]); b.size *= b.shape[i]; }
    if (b.size >= (1ull << 32)) return "output block with 2^32 or more elements";
    off += b.size;
    m.out_blocks.push_back(b);
    m.term_begin.push_back(uint32_t(m.terms.size()));
    m.terms.insert(m.terms.end(), kv.second.begin(), kv.second.end());
  }
  m.term_begin.push_back(uint32_t(m.terms.size()));
  m.out_elems = off;
  return "";
}

}  // namespace qlb200
