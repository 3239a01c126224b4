// plan.h -- execution context and per-contraction device plans (internal).
#ifndef QLB200_PLAN_H
#define QLB200_PLAN_H

#include <cstdint>
#include <string>
#include <vector>

#include "common.cuh"
#include "matcher.h"

struct qlb200_ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  // a contraction's few narrow-pair work items run beside its DMMA kernel: forked onto `side`, joined before returning
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // host <-> device pipelining (qlb200_hostpipe_*): copies run on their own stream beside the math
  cudaStream_t copy = nullptr;
  cudaStream_t alt = nullptr;     // odd-numbered parts of a split step run here, so that a part's tail overlaps the next part's head
  // grow-only arenas: `ws` holds the permuted operands, `stage` the device copies of host tensors
  void *ws = nullptr; size_t ws_bytes = 0;
  void *stage = nullptr; size_t stage_bytes = 0;
  uint64_t launches = 0;          // kernels launched by the last execute call
  uint64_t total_launches = 0;
  // captured graphs hold arena addresses: while any is alive, outgrown arenas are retired instead of freed
  int graphs_alive = 0;
  std::vector<void *> retired;
};

namespace qlb200 {

struct DeviceTables {
  PermBlk *perm_blks = nullptr;
  uint32_t *perm_tile_base = nullptr;
  GemmTask *tasks = nullptr;
  GemmGroup *groups = nullptr;
  GemmTile *tiles = nullptr;
  SkinnyItem *items = nullptr;
  uint32_t *seg = nullptr;
  unsigned int *counters = nullptr;
  void Free();
};

struct PlanHost {
  int dtype = 0;
  uint32_t flags = 0;
  int num_sms = 148;
  uint64_t a_elems = 0, b_elems = 0, c_elems = 0;      // raw sizes of the three tensors
  bool a_trans = false, b_trans = false;               // some block of A / B goes through the permute kernel
  uint64_t ws_a_elems = 0, ws_b_elems = 0;             // permuted operand sizes (workspace)
  std::vector<uint64_t> ws_off_a, ws_off_b;            // per block: element offset of its permuted copy, ~0 = read in place
  std::vector<PermBlk> perm_blks;
  std::vector<uint32_t> perm_tile_base;                // [nblk+1]
  // everything the whole contraction permutes; perm_blks / perm_tile_base above are the subset the current row partition needs
  std::vector<PermBlk> perm_blks_all;
  std::vector<uint32_t> perm_ntiles_all;
  std::vector<uint64_t> perm_owner_all;                // (operand << 63) | block ordinal
  std::vector<uint64_t> perm_size_all;
  std::vector<uint32_t> task_a_ord, task_b_ord;        // block ordinals of every GemmTask's operands
  std::vector<GemmTask> tasks;
  std::vector<GemmGroup> groups;                       // full row ranges
  std::vector<uint64_t> group_ksum;
  std::vector<GemmGroup> part_groups;                  // row ranges after partitioning
  std::vector<GemmTile> tiles;
  std::vector<uint32_t> seg;                           // stream-K: unit range of every CTA ([nseg + 1]); empty = dynamic
  uint32_t n_split_ctrs = 0;                           // split-K tiles (one arrival counter each)
  uint64_t n_part_slots = 0;                           // split-K partial-tile slots
  uint64_t part_slot_elems = 0;                        // elements per slot (BM x BN)
  std::vector<SkinnyItem> items;
  uint32_t skinny_sub = 1;                             // sub-chunks per narrow-pair item (see kSkinnyMaxSub)
  double flops = 0;
  uint64_t permute_elems_a = 0, permute_elems_b = 0;
  uint64_t gemm_read_bytes = 0, gemm_write_bytes = 0;
};

}  // namespace qlb200

struct qlb200_plan {
  qlb200_ctx *ctx = nullptr;
  qlb200::PlanHost h;
  qlb200::DeviceTables d;
  // parts of a split plan may run concurrently on two streams: each gets its own split-K partial-tile region of the arena
  uint64_t partials_shift = 0;    // bytes
  // accumulate form (qlb200_plan_create_accum): C_new = beta * C_old + alpha * sum of pairs
  bool accum = false;
  double alpha[2] = {1.0, 0.0}, beta[2] = {0.0, 0.0};
  uint64_t c_old_elems = 0;
  bool acc_expanded = false;                       // the output moves to a new (union) topology: C_old and C_new are distinct buffers
  std::vector<unsigned long long> acc_ranges;      // {old offset, new offset, length} of every untouched block that must move / scale
  unsigned long long *d_acc_ranges = nullptr;
};

struct qlb200_tplan {
  qlb200_ctx *ctx = nullptr;
  int dtype = 0;
  uint64_t elems = 0;
  int rank = 0;
  std::vector<uint64_t> blk_idx, offset;
  std::vector<uint32_t> coors, shape;
  std::vector<int8_t> scale;
  std::vector<qlb200::PermBlk> perm_blks;
  std::vector<uint32_t> perm_tile_base;
  qlb200::DeviceTables d;
};

namespace qlb200 {

/// Canonicalise one block permutation and choose its tiling. `perm[j]` = input axis of output axis j.
/// elem_bytes (8 / 16; 0 = unknown) lets run-mode blocks whose runs are 16-byte aligned take the bulk-copy path.
PermBlk MakePermBlk(int rank, const uint32_t *shape, const int32_t *perm, uint64_t src_off, uint64_t dst_off,
                    uint32_t src_sel, float scale, uint64_t *ntiles_out, int elem_bytes = 0);

/// Build the host-side tables of a contraction from sorted tasks.  `nctrct` = number of contracted
/// axes (the last nctrct entries of a_perm and the first nctrct of b_perm; 0 = unknown, keep order).
std::string BuildPlanHost(int dtype, uint32_t flags, int nctrct, int a_rank, const int32_t *a_perm,
                          uint64_t na, const uint32_t *a_shape, const uint64_t *a_off, uint64_t a_elems,
                          int b_rank, const int32_t *b_perm, uint64_t nb, const uint32_t *b_shape,
                          const uint64_t *b_off, uint64_t b_elems, const std::vector<qlb200_task> &sorted_tasks,
                          uint64_t c_elems, PlanHost *out);

/// (Re)build tile and item lists from part_groups.
std::string BuildTiles(PlanHost *h);

/// Restrict part_groups to the rows owned by `rank` of `world` (cost-balanced contiguous cut).
void PartitionRows(PlanHost *h, int world, int rank);

/// Keep only the permute descriptors of operand blocks that the current part_groups read (called by PartitionRows).
void FilterPermBlocks(PlanHost *h);

}  // namespace qlb200
#endif
