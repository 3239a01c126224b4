// matcher.h -- host-side quantum-number sector matcher (internal C++ interface behind qlb200_match_*).
// Replaces BlockSparseDataTensor::DataBlkGenForTenCtrct
// (reference: include/qlten/qltensor/blk_spar_data_ten/data_blk_operations.h:411-578).
#ifndef QLB200_MATCHER_H
#define QLB200_MATCHER_H

#include <cstdint>
#include <string>
#include <vector>

#include "qlb200.h"

namespace qlb200 {

/// Expanded copy of a qlb200_shell: per-block shapes, sizes, offsets and blk_idx precomputed.
struct Shell {
  int rank = 0;
  std::vector<uint32_t> nsct;
  std::vector<uint32_t> sct_base;   // start of index i inside deg/parity
  std::vector<uint32_t> deg;
  std::vector<uint8_t> parity;      // empty => bosonic
  std::vector<int8_t> dir;
  uint64_t nblk = 0;
  std::vector<uint32_t> coors;      // [nblk*rank]
  std::vector<uint32_t> shape;      // [nblk*rank]
  std::vector<uint64_t> size, offset, blk_idx;
  uint64_t elems = 0;

  bool fermionic() const { return !parity.empty(); }
  /// returns empty string on success, else the reason the shell is malformed
  std::string Load(const qlb200_shell *s);
};

struct CBlock {
  uint64_t blk_idx;
  uint64_t offset;
  uint64_t size;
  uint32_t coors[QLB200_MAX_RANK];
  uint32_t shape[QLB200_MAX_RANK];
};

struct Match {
  Shell a, b;                       // kept so plans can be built from the match alone
  std::vector<int> a_ctrct, b_ctrct, a_saved, b_saved;
  std::vector<int> a_perm, b_perm;  // saved_a ++ ctrct_a ; ctrct_b ++ saved_b
  bool a_need_trans = false, b_need_trans = false;
  int c_rank = 0;
  bool scalar = false;
  std::vector<uint32_t> c_nsct;
  std::vector<CBlock> c_blocks;     // ascending blk_idx
  uint64_t c_elems = 0;
  std::vector<qlb200_task> tasks;   // discovery order of the reference's (a, b) scan
  uint64_t candidate_pairs = 0;

  std::vector<qlb200_task> SortedTasks() const;
};

/// sel_axis < 0: all A blocks; otherwise only A blocks with coors[sel_axis] == sel_sector.
/// a_saved_order / b_saved_order (optional): the order in which the free axes of A / B appear in the result;
/// default is ascending (qlten::Contract).  mat_based_residue: multiply every task by the per-block signs of
/// BlockSparseDataTensor::CountResidueFermionSignForMatBasedCtrct (the contiguous-axes executor).
std::string BuildMatch(const qlb200_shell *a, const qlb200_shell *b, int nctrct, const int32_t *a_axes,
                       const int32_t *b_axes, int sel_axis, uint32_t sel_sector, Match *out,
                       const int32_t *a_saved_order = nullptr, const int32_t *b_saved_order = nullptr,
                       bool mat_based_residue = false);

/// Fermion exchange sign of one block pair (reference: data_blk_operations.h:351-401).
int FermionCtrctSign(const uint8_t *a_par, int a_rank, const uint8_t *b_par, int b_rank,
                     const std::vector<int> &a_ctrct, const std::vector<int> &b_ctrct,
                     const int8_t *a_dir);

/// Sign picked up by the odd-parity legs of a block when its legs are reordered
/// (reference: utility/utils_inl.h FermionicInplaceReorder, via DataBlk::Transpose data_blk.h:124-146).
int FermionReorderSign(const uint8_t *par, int rank, const int32_t *perm);

void EstimateCost(const Match &m, int dtype, qlb200_cost *out);

/// Output topology of the accumulate form C = beta * C + alpha * contract(A, B) (accum.cc).
struct AccumLayout {
  bool c_default = false;             // the output was a default tensor: the contraction's own topology
  bool expanded = false;              // required blocks were missing: the output is rebuilt on the union topology
  bool scalar = false;
  int rank = 0;
  std::vector<CBlock> blocks;         // resulting output blocks, ascending blk_idx, offsets in the NEW raw buffer
  std::vector<uint64_t> old_off;      // per resulting block: offset in the OLD raw buffer, ~0 = block is new
  std::vector<uint8_t> touched;       // per resulting block: the contraction writes it
  std::vector<uint64_t> req_to_union; // contraction-result block ordinal -> resulting block ordinal
  uint64_t elems = 0, old_elems = 0;  // raw sizes after / before
  qlb200_accum_stats stats{};
};
/// c_old == nullptr: default output.  Returns "" on success; *layout_mismatch tells a ContractAccumulateLayoutMismatch
/// (incompatible indexes / block shapes, or missing blocks with allow_expand == false) from a malformed argument.
std::string BuildAccumLayout(const Match &m, const qlb200_shell *c_old, bool c_old_has_data, bool allow_expand, int dtype, bool beta_zero,
                             bool beta_one, AccumLayout *out, bool *layout_mismatch);


/// Block pairing of the matrix-free axis operations (axis.cc): out = in with one / two axes multiplied by rank-2 operators.
struct AxisTerm { uint32_t in_ord, op1_ord, op2_ord; };
struct AxisMatch {
  Shell in, op[2];
  int nops = 0;
  int axis[2] = {-1, -1};
  std::vector<uint32_t> out_nsct;
  std::vector<CBlock> out_blocks;        // ascending blk_idx, offsets in the output raw buffer
  std::vector<uint32_t> term_begin;      // [nblk + 1] into terms
  std::vector<AxisTerm> terms;           // contributing (input block, op1 block, op2 block) per output block
  uint64_t out_elems = 0;
};
std::string BuildAxisMatch(const qlb200_shell *in, int nops, const qlb200_shell *op1, int axis1, const qlb200_shell *op2, int axis2,
                           AxisMatch *out);

}  // namespace qlb200
#endif
