// accum.cc -- output topology of the accumulate form  C = beta * C + alpha * contract(A, B)  (host only, no CUDA).
//
// Behavioural contract: MatrixBasedTensorContractionExecutor in accumulate mode,
// include/qlten/tensor_manipulation/contract_contiguous_axes.h -- GenerateDataBlk_ (:333-475, the accumulate branches
// :376-452), ExpandOutputTopology_ (:567-640), ScaleUntouchedOutputBlocks_ (:642-680), the first-task beta rule of
// ExecuteContractAccumulate_ (:716-731) and the counters of ContiguousContractStats (:55-80).  What the reference does
// with tensor objects and per-block BLAS calls is flattened here into one table: for every block of the resulting
// output (the union of the existing blocks and the blocks the contraction produces) where it lies in the new raw
// buffer, where it lay in the old one, and whether the contraction touches it.
#include "matcher.h"

#include <algorithm>
#include <cstring>

namespace qlb200 {

std::string BuildAccumLayout(const Match &m, const qlb200_shell *c_old, bool c_old_has_data, bool allow_expand, int dtype, bool beta_zero, bool beta_one,
                             AccumLayout *out, bool *layout_mismatch) {
  AccumLayout &L = *out;
  *layout_mismatch = false;
  const uint64_t es = dtype == QLB200_C64 ? 16 : 8;
  L = AccumLayout();
  L.scalar = m.scalar;
  L.rank = m.c_rank;
  qlb200_accum_stats &st = L.stats;
  st.raw_data_contract_tasks = st.gemm_calls = st.accumulate_gemm_calls = m.tasks.size();
  st.accumulate_calls = 1;
  uint64_t required_bytes = 0;
  for (const CBlock &cb : m.c_blocks) required_bytes += cb.size * es;
  if (m.scalar && !m.tasks.empty()) required_bytes = es;

  if (c_old == nullptr) {
    // default output: the contraction's own topology, every block new (beta must be 0; checked by the caller)
    L.c_default = true;
    L.blocks = m.c_blocks;
    L.elems = m.c_elems;
    L.old_off.assign(L.blocks.size(), ~0ull);
    L.touched.assign(L.blocks.size(), 1);
    L.req_to_union.resize(m.c_blocks.size());
    for (size_t i = 0; i < m.c_blocks.size(); ++i) L.req_to_union[i] = i;
    st.output_tensor_rebuilds = 1;
    st.temporary_output_bytes_avoided = required_bytes;
    return "";
  }

  Shell old;
  std::string err = old.Load(c_old);
  if (!err.empty()) return "C: " + err;
  // the indexes of an existing output must be those of the contraction result (the adapter compares the Index objects;
  // here: rank, sector counts and degeneracies)
  if (old.rank != m.c_rank) { *layout_mismatch = true; return "output rank differs from the contraction result"; }
  {
    size_t pos = 0;
    for (int i = 0; i < m.c_rank; ++i) {
      if (old.nsct[i] != m.c_nsct[i]) { *layout_mismatch = true; return "output index has a different sector count"; }
      pos += old.nsct[i];
    }
    (void) pos;
  }
  L.old_elems = old.elems;
  st.temporary_output_bytes_avoided = required_bytes;
  if (m.scalar) {
    // rank-0 output: one element, every task accumulates into it; with no task the value is an untouched "block"
    L.elems = 1;
    L.old_elems = c_old_has_data ? 1 : 0;
    if (m.tasks.empty() && !beta_one) st.output_untouched_scale_bytes += es;    // Fill(0) gives an unallocated scalar its element first (:691-693)
    return "";
  }

  // union of the old blocks and the required ones, ascending blk_idx (std::map order)
  size_t io = 0, ir = 0;
  bool missing = false;
  L.req_to_union.resize(m.c_blocks.size());
  while (io < old.nblk || ir < m.c_blocks.size()) {
    const bool take_old = ir >= m.c_blocks.size() || (io < old.nblk && old.blk_idx[io] <= m.c_blocks[ir].blk_idx);
    const bool take_req = io >= old.nblk || (ir < m.c_blocks.size() && m.c_blocks[ir].blk_idx <= old.blk_idx[io]);
    CBlock b;
    std::memset(&b, 0, sizeof(b));
    if (take_req) {
      b = m.c_blocks[ir];
      if (take_old) {     // present on both sides: shapes must agree
        for (int i = 0; i < m.c_rank; ++i)
          if (old.shape[io * old.rank + i] != b.shape[i] || old.coors[io * old.rank + i] != b.coors[i]) {
            *layout_mismatch = true;
            return "output block shape is not compatible with the contraction result";
          }
      }
    } else {
      b.blk_idx = old.blk_idx[io]; b.size = old.size[io];
      for (int i = 0; i < old.rank; ++i) { b.coors[i] = old.coors[io * old.rank + i]; b.shape[i] = old.shape[io * old.rank + i]; }
    }
    L.old_off.push_back(take_old ? old.offset[io] : ~0ull);
    L.touched.push_back(take_req ? 1 : 0);
    if (take_req) L.req_to_union[ir] = L.blocks.size();
    if (!take_old) missing = true;
    L.blocks.push_back(b);
    if (take_old) ++io;
    if (take_req) ++ir;
  }
  if (missing && !allow_expand) {
    *layout_mismatch = true;
    return "output block topology requires expansion";
  }
  uint64_t off = 0;
  for (CBlock &b : L.blocks) { b.offset = off; off += b.size; }
  L.elems = off;
  L.expanded = missing;
  if (L.expanded) {
    // ExpandOutputTopology_: every old block is copied (or scale-copied when untouched and beta != 1); beta == 0 copies nothing
    st.output_topology_expansions = 1;
    st.output_tensor_rebuilds = 1;
    for (size_t u = 0; u < L.blocks.size(); ++u) {
      if (L.old_off[u] == ~0ull) { ++st.output_expand_new_blocks; continue; }
      if (beta_zero) continue;
      st.output_expand_copy_bytes += L.blocks[u].size * es;
      if (!L.touched[u] && !beta_one) st.output_untouched_scale_bytes += L.blocks[u].size * es;
    }
  } else if (!beta_one) {
    // ScaleUntouchedOutputBlocks_
    for (size_t u = 0; u < L.blocks.size(); ++u)
      if (!L.touched[u]) st.output_untouched_scale_bytes += L.blocks[u].size * es;
  }
  return "";
}

}  // namespace qlb200
