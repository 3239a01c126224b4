/* qlb200.h -- C ABI of the B200-native block-sparse contraction path.
 *
 * Drop-in boundary for ONE path of QuantumLiquids/TensorToolkit: qlten::Contract on
 * block-sparse symmetric tensors.  The reference has no FFI of its own (header-only C++
 * templates, backends chosen by macros), so the entry points below are cut at the seam the
 * reference itself draws between "block metadata" and "raw data":
 *
 *   qlb200_match            replaces BlockSparseDataTensor::DataBlkGenForTenCtrct
 *                           (include/qlten/qltensor/blk_spar_data_ten/data_blk_operations.h:411-578)
 *                           + TenCtrctGenSavedAxesSet (:268-296) + TenCtrctNeedTransCheck
 *                           (include/qlten/tensor_manipulation/ten_ctrct.h:396-443)
 *                           + FermionExchangeSignForCtrct (data_blk_operations.h:351-401)
 *                           + DataBlksOffsetRefresh (:137-145)
 *   qlb200_task (struct)    replaces RawDataCtrctTask
 *                           (include/qlten/qltensor/blk_spar_data_ten/raw_data_operation_tasks.h:196-270)
 *   qlb200_plan_* / qlb200_execute
 *                           replace BlockSparseDataTensor::CtrctTwoBSDTAndAssignIn
 *                           (include/qlten/qltensor/blk_spar_data_ten/global_operations.h:895-992),
 *                           i.e. every hp_numeric::TensorTranspose (framework/hp_numeric/ten_trans.h:94-114
 *                           HPTT, :245-349 cuTENSOR) and hp_numeric::MatMultiply
 *                           (framework/hp_numeric/blas_level3.h:35-108 CBLAS, :798-981 cuBLAS) call of a
 *                           contraction, in two kernel launches (batched permute + grouped GEMM).
 *   qlb200_transpose_*      replaces BlockSparseDataTensor::Transpose raw-data part
 *                           (global_operations.h:393-441, raw_data_operations.h:201-218).
 *   qlb200_estimate_cost    replaces EstimateContractCost
 *                           (include/qlten/tensor_manipulation/tensor_op_cost.h:467-516).
 *
 * Plain C types only; every array argument is caller-owned and read during the call only.
 * All functions return QLB200_OK (0) or a negative error code; qlb200_last_error() gives text.
 * There is NO CPU fallback: execute functions fail with QLB200_ERR_CUDA when no sm_100 device
 * is usable.
 */
#ifndef QLB200_H
#define QLB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QLB200_MAX_RANK 8

enum {
  QLB200_OK = 0,
  QLB200_ERR_ARG = -1,      /* precondition violated (the reference only assert()s these) */
  QLB200_ERR_CUDA = -2,     /* CUDA runtime error / no device */
  QLB200_ERR_NOMEM = -3,
  QLB200_ERR_UNSUPPORTED = -4,
  QLB200_ERR_LAYOUT = -5    /* accumulate form: the existing output cannot take the result (the reference's
                               detail::ContractAccumulateLayoutMismatch; Try... returns false on it) */
};

enum { QLB200_F64 = 0, QLB200_C64 = 1 };               /* double, complex double (interleaved re,im) */
enum { QLB200_MEM_HOST = 0, QLB200_MEM_DEVICE = 1 };   /* where A/B/C raw buffers live */
enum { QLB200_DIR_IN = -1, QLB200_DIR_OUT = 1 };       /* TenIndexDirType, qltensor/index.h:34-38 */

/* Block "shell" of one BlockSparseDataTensor: everything the matcher reads, nothing it does not.
 * Blocks are listed in ascending blk_idx order (std::map order, blk_spar_data_ten.h:459) where
 * blk_idx = row-major index of blk_coors over nsct (blk_spar_data_ten.h:380-382); data offsets are
 * the prefix sums of block sizes in that order (data_blk_operations.h:137-145). */
typedef struct qlb200_shell {
  int32_t rank;              /* number of indexes, 1..QLB200_MAX_RANK */
  const uint32_t *nsct;      /* [rank]   sectors per index (BSDT blk_shape) */
  const uint32_t *deg;       /* [sum nsct] degeneracy of every sector, index-major */
  const uint8_t *parity;     /* [sum nsct] 1 = odd fermion parity; NULL for bosonic QN types */
  const int8_t *dir;         /* [rank]   QLB200_DIR_IN / QLB200_DIR_OUT (only read when parity!=NULL) */
  uint64_t nblk;             /* stored blocks */
  const uint32_t *blk_coors; /* [nblk*rank] sector coordinates of each stored block */
} qlb200_shell;

/* One matched block pair == one GEMM.  Field meaning follows RawDataCtrctTask. Offsets are in
 * elements into the raw buffers of A, B and C. */
typedef struct qlb200_task {
  uint64_t a_blk_idx, b_blk_idx, c_blk_idx;
  uint64_t a_off, b_off, c_off;
  uint32_t a_ord, b_ord, c_ord; /* ordinal of the block in its tensor's ascending-blk_idx list */
  uint32_t m, k, n;
  int8_t sign;                  /* fermion exchange sign, +1 / -1 (f_ex_sign) */
  uint8_t first;                /* 1 = this pair creates the C block (beta = 0) */
  uint8_t pad_[2];
} qlb200_task;

typedef struct qlb200_cost {    /* same accounting as tensor_op_cost.h:109-120 */
  double flops;
  uint64_t gemm_count;
  uint64_t candidate_block_pair_count;
  uint64_t output_block_count;
  uint64_t output_raw_elem_count;
  uint64_t read_bytes, write_bytes, temp_peak_bytes;
} qlb200_cost;

typedef struct qlb200_match qlb200_match;   /* result of the sector matcher (host only) */
typedef struct qlb200_ctx qlb200_ctx;       /* one per device / host thread */
typedef struct qlb200_plan qlb200_plan;     /* device-side descriptor tables of one contraction */
typedef struct qlb200_tplan qlb200_tplan;   /* device-side descriptor table of a whole-tensor transpose */

/* ---- library ------------------------------------------------------------------------------ */
const char *qlb200_version(void);
const char *qlb200_last_error(void);        /* thread-local text of the last failure */

/* ---- sector matcher (host, no CUDA calls) --------------------------------------------------- */
int qlb200_match_create(const qlb200_shell *a, const qlb200_shell *b, int32_t nctrct,
                        const int32_t *a_axes, const int32_t *b_axes, qlb200_match **out);
void qlb200_match_destroy(qlb200_match *m);
/* Restrict matching to A blocks whose coordinate on `axis` equals `sector`: the block set of
 * dmrg::Contract1Sector (tensor_manipulation/dmrg/contract_1sector.h:181-228). */
int qlb200_match_create_1sector(const qlb200_shell *a, int32_t axis, uint32_t sector,
                                const qlb200_shell *b, int32_t nctrct, const int32_t *a_axes,
                                const int32_t *b_axes, qlb200_match **out);
/* Contiguous-axes contraction: the block pairing, result layout and fermion signs of
 * qlten::ContractContiguousAxes / MatrixBasedTensorContractionExecutor
 * (tensor_manipulation/contract_contiguous_axes.h:205-331, 333-364, 849-873).  The contracted axes are
 * (a_start + i) % rank_a and (b_start + i) % rank_b, i < size; the free axes of each operand enter the result in
 * CYCLIC order starting behind the contracted range.  The plan built from this match reads every operand block
 * in place -- a cyclic rotation is one 2-D transposition of the block, absorbed by the GEMM's operand loads --
 * so the reference's OutOfPlaceMatrixTransposeForSelectedDataBlk pass has no counterpart here.  The result is the
 * same for every CtrctSide pair (they are performance hints in the reference). */
int qlb200_match_create_contiguous(const qlb200_shell *a, const qlb200_shell *b, int32_t a_start, int32_t b_start,
                                   int32_t size, qlb200_match **out);
/* the free axes of A (which = 0) / B (1) in result order; returns how many */
int32_t qlb200_match_saved_axes(const qlb200_match *m, int which, int32_t *axes_out);
int32_t qlb200_match_c_rank(const qlb200_match *m);
uint64_t qlb200_match_c_nblk(const qlb200_match *m);
uint64_t qlb200_match_c_elems(const qlb200_match *m);            /* raw_data_size_ of C */
uint64_t qlb200_match_ntask(const qlb200_match *m);
int qlb200_match_is_scalar(const qlb200_match *m);
/* perm = saved_a ++ ctrct_a (A) / ctrct_b ++ saved_b (B); returns 1 if a transpose is needed */
int qlb200_match_perm(const qlb200_match *m, int which /*0=A,1=B*/, int32_t *perm_out);
/* C blocks in ascending blk_idx order */
int qlb200_match_c_blocks(const qlb200_match *m, uint64_t *blk_idx, uint32_t *blk_coors,
                          uint32_t *shape, uint64_t *offset);
/* tasks: order==0 -> discovery order of the reference's (a,b) scan; order==1 -> sorted by
 * (c_blk_idx, first-task-first), the order SortTasksByCBlkIdx establishes (stable here). */
int qlb200_match_tasks(const qlb200_match *m, int order, qlb200_task *tasks_out);
int qlb200_estimate_cost(const qlb200_match *m, int dtype, qlb200_cost *out);

/* ---- execution ------------------------------------------------------------------------------ */
int qlb200_ctx_create(int device, qlb200_ctx **out);
void qlb200_ctx_destroy(qlb200_ctx *ctx);
int qlb200_ctx_sync(qlb200_ctx *ctx);
void *qlb200_ctx_stream(qlb200_ctx *ctx);                         /* cudaStream_t */
int qlb200_ctx_set_stream(qlb200_ctx *ctx, void *cuda_stream);    /* run on a caller stream */

/* device memory helpers for host languages without a CUDA binding */
int qlb200_dev_alloc(qlb200_ctx *ctx, size_t bytes, void **out);
int qlb200_dev_free(qlb200_ctx *ctx, void *p);
int qlb200_memcpy_h2d(qlb200_ctx *ctx, void *dst, const void *src, size_t bytes);  /* async on ctx stream */
int qlb200_memcpy_d2h(qlb200_ctx *ctx, void *dst, const void *src, size_t bytes);  /* async on ctx stream */
int qlb200_host_register(void *p, size_t bytes);                  /* pin caller memory */
int qlb200_host_unregister(void *p);

/* Build the device descriptor tables from a match.  flags: see below.
 * ctx == NULL builds a host-only plan: stats, partition and c_ranges work, execute does not. */
#define QLB200_PLAN_DETERMINISTIC 1u   /* default and only mode: no atomics, fixed summation order */
#define QLB200_PLAN_NO_SKINNY 2u       /* force every task through the DMMA kernel (testing) */
/* 4u: reserved (the first-generation cp.async kernels it selected were removed in round 2) */
#define QLB200_PLAN_CPLX_4M 32u       /* complex GEMM: four real DMMAs per complex step instead of the default three
                                         (Gauss / 3M product: 25 % fewer tensor-pipe instructions, same error order) */
#define QLB200_PLAN_STAGGER_OUTPUT 64u /* cut every long k loop into ~4 units queued back to back, so that output tiles complete
                                         throughout the launch instead of all at its end: lets a fused multi-GPU exchange
                                         (execute_bcast / execute_mcast) overlap the NVLink transfer with the remaining math */
#define QLB200_PLAN_STREAM_K 128u      /* stream-K schedule: the tiles' k loops, laid end to end, are cut into one equal-cost
                                         segment per resident CTA (static assignment); a tile that straddles a cut is split
                                         and finished by the deterministic split-K fix-up.  Opt-in experiment: balances to within 2 % on paper
                                         but measured SLOWER than the default weighted-LPT list on the 8-way shards of the headline
                                         workload (0.60 / 0.63 ms vs 0.54 / 0.57 ms per GEMM step): ~2 partial tiles per CTA cost
                                         more than the imbalance they remove. */
#define QLB200_PLAN_NO_SPLIT_K 16u     /* never cut a tile's k loop into several units (testing / tuning) */
#define QLB200_PLAN_NO_VIEW 256u       /* do not read (n1, k, n2)-stored B blocks as strided k x (n1 n2) views in the GEMM producer: send them
                                          through the permute kernel instead (testing / measuring the permute kernel) */
#define QLB200_PLAN_PERMUTE_ALL 8u     /* send every block of a transposed operand through the permute kernel
                                          (default: blocks whose permutation is trivial or one 2-D transposition
                                          are read in place by the GEMM) */
int qlb200_plan_create(qlb200_ctx *ctx, const qlb200_match *m, const qlb200_shell *a,
                       const qlb200_shell *b, int dtype, uint32_t flags, qlb200_plan **out);
/* Descriptor-table entry (no shells): permute every A/B block with one perm each, then run the
 * grouped GEMM over `tasks`.  a_shape/b_shape are [n*rank] block shapes, offsets in elements. */
int qlb200_plan_create_raw(qlb200_ctx *ctx, int dtype, uint32_t flags, int32_t a_rank,
                           const int32_t *a_perm, uint64_t na, const uint32_t *a_shape,
                           const uint64_t *a_off, int32_t b_rank, const int32_t *b_perm, uint64_t nb,
                           const uint32_t *b_shape, const uint64_t *b_off, uint64_t ntask,
                           const qlb200_task *tasks, uint64_t c_elems, qlb200_plan **out);
void qlb200_plan_destroy(qlb200_plan *p);
/* Keep only the output row slabs of this rank (multi-GPU: partition by output sector / row slab).
 * Units are whole C blocks or row ranges of C blocks, balanced by flops (LPT). */
int qlb200_plan_partition(qlb200_plan *p, int32_t world, int32_t rank);
/* ranges of C (element offsets, lengths) this plan writes after partitioning */
uint64_t qlb200_plan_c_range_count(const qlb200_plan *p);
int qlb200_plan_c_ranges(const qlb200_plan *p, uint64_t *off, uint64_t *len);

typedef struct qlb200_plan_stats {
  double flops;                 /* sum over tasks, 2mkn or 8mkn */
  uint64_t ntask, ngroup, ntile_dmma, nrow_skinny;
  uint64_t permute_elems_a, permute_elems_b;   /* elements moved by the batched permute */
  uint64_t workspace_bytes;
  uint64_t gemm_read_bytes, gemm_write_bytes;  /* (mk+kn)*s per task, mn*s per C block */
} qlb200_plan_stats;
int qlb200_plan_get_stats(const qlb200_plan *p, qlb200_plan_stats *out);
/* Where the GEMM reads block `ord` of operand `which` (0 = A, 1 = B): returns 0 when it is read in place from the
 * caller's buffer (identity or one 2-D transposition), 1 when it goes through the batched permute kernel -- *ws_off is
 * then the element offset of its permuted (row-major m x k / k x n) copy inside that operand's workspace region --
 * or a negative error code. */
int qlb200_plan_operand_block(const qlb200_plan *p, int which, uint64_t ord, uint64_t *ws_off);
/* Copy `elems` elements at `elem_off` of operand `which`'s permuted workspace to host memory (after
 * qlb200_execute_permute or qlb200_execute; synchronises).  Lets a test compare the permute kernel's output bit for bit
 * with the reference's per-block hp_numeric::TensorTranspose (ten_trans.h:94-114). */
int qlb200_plan_read_workspace(qlb200_ctx *ctx, qlb200_plan *p, int which, uint64_t elem_off, uint64_t elems, void *dst_host);
/* The DMMA work-unit list in launch order (introspection for tests / tuning): a unit is the k-stage range
 * [s_begin, s_end) of output tile (tm, tn) of one output block; a tile cut along k has nsplit units (split = its slot).
 * Geometry of the plan's kernel: tile rows x cols and k elements per stage.  Returns the number of units; fills at
 * most `cap` entries. */
typedef struct qlb200_unit {
  uint32_t group, tm, tn;          /* output block (ordinal among the plan's output blocks), tile coordinates */
  uint32_t s_begin, s_end;         /* k-stage range */
  uint32_t split, nsplit;
  uint32_t rows, cols;             /* valid extent of the tile inside the (partitioned) block */
} qlb200_unit;
uint64_t qlb200_plan_units(const qlb200_plan *p, uint64_t cap, qlb200_unit *out, uint32_t *tile_rows, uint32_t *tile_cols,
                           uint32_t *stage_k);
/* Work items of the narrow-pair kernel (output blocks with n <= 8 and every k <= 32): item i covers rows
 * [row0, row0 + rows) of output block `group`.  Returns their number; fills at most `cap`. */
typedef struct qlb200_item {
  uint32_t group, row0, rows;
  uint32_t n;                      /* columns of the output block */
} qlb200_item;
uint64_t qlb200_plan_items(const qlb200_plan *p, uint64_t cap, qlb200_item *out);
/* QLB200_PLAN_STREAM_K plans: CTA b runs units [seg[b], seg[b+1]).  Returns the number of table entries (CTAs + 1; 0 for
 * plans whose units are pulled dynamically); fills at most `cap`. */
uint64_t qlb200_plan_segments(const qlb200_plan *p, uint64_t cap, uint32_t *seg_out);

/* C = contract(A, B).  C must hold c_elems elements (uninitialised is fine).  With
 * QLB200_MEM_HOST the call stages A, B through device memory and copies C back, then
 * synchronises; with QLB200_MEM_DEVICE it only enqueues work on the ctx stream. */
int qlb200_execute(qlb200_ctx *ctx, qlb200_plan *p, const void *A, const void *B, void *C,
                   int mem_kind);
/* the two phases separately (device pointers only) -- used by the benchmark to time each kernel */
int qlb200_execute_permute(qlb200_ctx *ctx, qlb200_plan *p, const void *A, const void *B);
int qlb200_execute_gemm(qlb200_ctx *ctx, qlb200_plan *p, const void *A, const void *B, void *C);
/* ---- host <-> device pipelining of a contraction chain ------------------------------------------- */
/* Split a plan into `nparts` plans that together compute exactly what `p` computes (disjoint sets of output blocks), so
 * that a PCIe transfer can overlap the math:
 *   QLB200_SPLIT_BY_A / _BY_B  part i reads only elements of operand A / B below bounds_out[i + 1]: the operand can be
 *                              streamed to the device in nparts chunks, part i starting as soon as chunk i has landed;
 *   QLB200_SPLIT_BY_C          part i writes exactly the range [bounds_out[i], bounds_out[i + 1]) of C: the result can be
 *                              streamed back, chunk i leaving while part i + 1 computes.
 * cum_frac[i] = wanted cumulative share of the operand's elements at the end of chunk i (ascending, last = 1); cuts snap
 * to block boundaries.  Plans with blocks that need the permute kernel are not split (QLB200_ERR_UNSUPPORTED). */
enum { QLB200_SPLIT_BY_A = 0, QLB200_SPLIT_BY_B = 1, QLB200_SPLIT_BY_C = 2 };
int qlb200_plan_split(const qlb200_plan *p, int by, int32_t nparts, const double *cum_frac, qlb200_plan **parts_out,
                      uint64_t *bounds_out /* [nparts + 1], elements */);
/* End-to-end apply of a chain whose input operand lives in HOST memory and whose result goes back to HOST memory
 * (a Lanczos mat-vec driven from the host: the reference's Contract chain on host tensors).  The first step's streamed
 * operand is uploaded in chunks on a copy stream while the parts of that step run; the last step's parts are followed by
 * the download of their output ranges.  Everything else of the chain is enqueued by the caller between begin and end. */
typedef struct qlb200_hostpipe qlb200_hostpipe;
int qlb200_hostpipe_create(qlb200_ctx *ctx, const qlb200_plan *first, int first_streams /* QLB200_SPLIT_BY_A or _BY_B */,
                           int32_t nparts_in, const double *cum_in, const qlb200_plan *last, int32_t nparts_out,
                           const double *cum_out, qlb200_hostpipe **out);
void qlb200_hostpipe_destroy(qlb200_hostpipe *hp);
/* in_host (pinned for real overlap) -> in_dev in chunks; first step: C = contract(A, B) where the streamed operand is
 * in_dev and `other_dev` is the resident one.  Enqueues only. */
int qlb200_hostpipe_begin(qlb200_ctx *ctx, qlb200_hostpipe *hp, const void *in_host, void *in_dev, const void *other_dev, void *c_dev);
/* last step on device operands, result streamed to out_host; returns after the last byte has arrived (synchronises). */
int qlb200_hostpipe_end(qlb200_ctx *ctx, qlb200_hostpipe *hp, const void *a_dev, const void *b_dev, void *c_dev, void *out_host);
uint64_t qlb200_hostpipe_launches(const qlb200_hostpipe *hp);   /* kernels launched by the last begin + end */

/* ---- accumulate form: C = beta * C + alpha * contract(A, B) ------------------------------------ */
/* Replaces the accumulate mode of MatrixBasedTensorContractionExecutor behind qlten::ContractTailHeadContiguousAccumulate /
 * TryContractTailHeadContiguousAccumulate (tensor_manipulation/contract_contiguous_axes.h:954-1041; executor :333-475,
 * :567-782).  The existing output may hold MORE blocks than the contraction produces (they are only scaled by beta) or
 * FEWER (allow_expand: the output is rebuilt on the union topology; new blocks get a first-task beta of zero).  Everything
 * the reference does with per-block VectorCopy / VectorScale / GEMM-with-beta calls happens in at most three launches:
 * batched permute (if any block needs it), one scale-copy launch over the untouched blocks, and the grouped GEMM whose
 * epilogue reads the old block once and writes  beta * C_old + alpha * sum(pairs)  at the block's new place. */
typedef struct qlb200_accum_stats {   /* the ContiguousContractStats counters (contract_contiguous_axes.h:55-80) this path defines */
  uint64_t raw_data_contract_tasks, gemm_calls, accumulate_calls, accumulate_gemm_calls;
  uint64_t output_tensor_rebuilds, temporary_output_bytes_avoided;
  uint64_t output_topology_expansions, output_expand_copy_bytes, output_expand_new_blocks, output_untouched_scale_bytes;
} qlb200_accum_stats;
typedef struct qlb200_accum qlb200_accum;   /* resulting output topology of one accumulate call (host only) */
/* c_old: shell of the existing output, NULL for a default tensor (then beta must be 0: QLB200_ERR_ARG otherwise).
 * c_old_has_data: its raw buffer is allocated.  alpha / beta: (re, im); im ignored for QLB200_F64.
 * Returns QLB200_ERR_LAYOUT when the indexes / block shapes are incompatible or blocks are missing and !allow_expand. */
int qlb200_accum_create(const qlb200_match *m, const qlb200_shell *c_old, int c_old_has_data, int allow_expand, int dtype,
                        const double *alpha2, const double *beta2, qlb200_accum **out);
void qlb200_accum_destroy(qlb200_accum *a);
uint64_t qlb200_accum_nblk(const qlb200_accum *a);
uint64_t qlb200_accum_elems(const qlb200_accum *a);        /* raw size of the resulting output */
int qlb200_accum_expanded(const qlb200_accum *a);          /* 1: the output must be rebuilt on the union topology */
/* resulting blocks in ascending blk_idx order; old_offset = ~0 for blocks the contraction adds; touched = the contraction writes it */
int qlb200_accum_blocks(const qlb200_accum *a, uint64_t *blk_idx, uint32_t *blk_coors, uint32_t *shape, uint64_t *offset,
                        uint64_t *old_offset, uint8_t *touched);
int qlb200_accum_get_stats(const qlb200_accum *a, qlb200_accum_stats *out);
int qlb200_plan_create_accum(qlb200_ctx *ctx, const qlb200_match *m, const qlb200_accum *a, int dtype, uint32_t flags,
                             qlb200_plan **out);
/* C_new = beta * C_old + alpha * contract(A, B) on the resulting topology.  C_old may be NULL when beta == 0 or the output
 * was default.  Device pointers: same topology (qlb200_accum_expanded == 0) runs IN PLACE, C_old == C_new; an expanded
 * topology needs a new buffer of qlb200_accum_elems elements, C_new != C_old.  Host pointers: any combination. */
int qlb200_execute_accum(qlb200_ctx *ctx, qlb200_plan *p, const void *A, const void *B, const void *C_old, void *C_new,
                         int mem_kind);

/* ---- multi-GPU: exchange fused into the GEMM epilogue ---------------------------------------- */
/* Ranks that own disjoint row slabs of one result (qlb200_plan_partition, or operands restricted to
 * a row range) can write them straight into the full result on EVERY GPU of the NVLink domain: the
 * output tiles of the grouped GEMM are stored to all `npeers` buffers (the caller's own and its peers',
 * mapped with qlb200_ipc_open) from inside the kernel, so no separate all-gather is needed -- only a
 * barrier before the result is read.  Device pointers only; npeers <= 8.
 * ORDERING IN A LOOP: the barrier that follows the stores orders them before the peers' reads of THIS result; nothing
 * orders the NEXT call's stores after a slower peer's reads.  When the result of call j is the input of call j+1
 * (Lanczos), alternate between two result buffers (call j writes buffer j & 1, as tensortoolkit_b200.ShardedChain and
 * qlten::b200 do) or put a second barrier in front of the storing call. */
int qlb200_execute_bcast(qlb200_ctx *ctx, qlb200_plan *p, const void *A, const void *B, void *const *C_peers,
                         int32_t npeers);
/* Same exchange through the NVSwitch: `C_multicast` is a multicast (NVLS) mapping of the full result buffer of
 * every GPU (cuMulticastCreate / cuMulticastBindMem, or torch.distributed._symmetric_memory's multicast_ptr).
 * Each output tile leaves the GPU once as multimem.st stores and the switch writes it into every replica, the
 * caller's own included -- 1/npeers of the NVLink traffic of qlb200_execute_bcast.  A barrier is still needed
 * before the result is read. */
int qlb200_execute_mcast(qlb200_ctx *ctx, qlb200_plan *p, const void *A, const void *B, void *C_multicast);
/* Fan one GPU's share of a replicated buffer out to every GPU: bytes [byte_off, byte_off + bytes) of `src` (this GPU's
 * replica, just uploaded from the host) are stored at the same offset of every peer buffer (npeers unicast pointers) or of
 * the NVSwitch multicast mapping `dst_multicast` (one multimem.st per 16 bytes, replicated by the switch).  With N ranks
 * each uploading 1/N of the next input over its own PCIe link, the whole input is on every GPU after one barrier.
 * Offsets / lengths in bytes, multiples of 16.  Give either dst_peers or dst_multicast. */
int qlb200_fanout_copy(qlb200_ctx *ctx, const void *src, uint64_t byte_off, uint64_t bytes, void *const *dst_peers, int32_t npeers,
                       void *dst_multicast);
/* Rebase output blocks: the block the plan would write at element offset from_off[i] is written at
 * to_off[i] instead (a rank's packed row slabs -> their place in the full result layout). */
int qlb200_plan_remap_output(qlb200_plan *p, uint64_t n, const uint64_t *from_off, const uint64_t *to_off);
/* CUDA IPC plumbing for host languages without a CUDA binding: export a 64-byte handle of a buffer
 * obtained from qlb200_dev_alloc, open a peer's handle (enables peer access), close it. */
int qlb200_ipc_export(qlb200_ctx *ctx, const void *dev_ptr, unsigned char *handle64);
int qlb200_ipc_open(qlb200_ctx *ctx, const unsigned char *handle64, void **peer_ptr);
int qlb200_ipc_close(qlb200_ctx *ctx, void *peer_ptr);

/* ---- row-slab partitioner of a contraction chain (host only) ------------------------------------------------------- */
/* The reference distributes the DMRG mat-vec over MPI ranks by restricting one FREE index of the first operand to one QN
 * sector per work unit (dmrg::Contract1Sector, dmrg/contract_1sector.h:181-228).  Here the unit is a range of ROWS of a
 * sector: the rows of the split index form one line (sector-major) of weighted pieces, rank r owns the r-th equal-weight
 * segment.  All ranks compute the same cuts from the same numbers.  (tensortoolkit_b200/sharding.py is the Python face.) */
typedef struct qlb200_piece {
  uint32_t sector, lo, hi;         /* rows [lo, hi) of one sector of the split index */
  uint32_t pad_;
  double weight;                   /* cost per row */
} qlb200_piece;
/* cost[s] += flops of every matched pair whose A block lies in sector s of A's index `axis` (2 m k n, 8 m k n complex):
 * summed over the steps of a chain (the split index stays a free index of the left operand) this is the line's weight */
int qlb200_shard_sector_flops(const qlb200_match *m, int32_t axis, int dtype, double *cost);
/* ranges_out[(r * nsct + s) * 2 + {0, 1}] = rows [lo, hi) of sector s owned by rank r; cuts inside a sector are multiples of
 * `snap` rows.  A rank may come out empty (world larger than what can be cut). */
int qlb200_shard_cut_line(const qlb200_piece *pieces, uint64_t npieces, const uint32_t *degs, uint32_t nsct, int32_t world,
                          int32_t snap, uint32_t *ranges_out);
/* Feedback step: times[r] = measured time of rank r's share under `ranges` (cut from `pieces`); every rank's rows are
 * re-weighted by ((time share) / (modelled share)) ^ damp and returned as pieces split at the old cuts (sorted by sector,
 * row).  Returns their number (every old piece appears once per rank that owns rows of it: at most npieces + world - 1 for
 * cuts made from the same pieces); fills at most `cap` -- call with cap = 0 first. */
uint64_t qlb200_shard_reweigh(const qlb200_piece *pieces, uint64_t npieces, const uint32_t *ranges, uint32_t nsct, int32_t world,
                              const double *times, double damp, uint64_t cap, qlb200_piece *out);

/* Row slab of a tensor: keep rows [ranges[2 s], ranges[2 s + 1]) of every sector s of index `axis`; sectors left empty
 * disappear, and their blocks with them (the restricted operand of a rank; the reference restricts to ONE sector,
 * dmrg/contract_1sector.h:181-228).  Call once with the output arrays NULL for the counts in *info, then with arrays of those
 * sizes.  kept_sectors[i] = old number of the slab's sector i (new_deg[i] its degeneracy); kept_blocks[j] = old ordinal of
 * the slab's block j, new_coors[j * rank ..] its coordinates in the slab; (copy_src, copy_dst, copy_len)[k] = element ranges
 * that build the slab's raw buffer from the full one (blocks packed in order) -- directly, or on the device with
 * qlb200_cplan_create. */
typedef struct qlb200_slab_info {
  uint32_t nsct_kept;
  uint32_t pad_;
  uint64_t nblk_kept, elems, ncopy;
} qlb200_slab_info;
int qlb200_shard_restrict(const qlb200_shell *t, int32_t axis, const uint32_t *ranges, qlb200_slab_info *info,
                          uint32_t *kept_sectors, uint32_t *new_deg, uint32_t *kept_blocks, uint32_t *new_coors,
                          uint64_t *copy_src, uint64_t *copy_dst, uint64_t *copy_len);

/* ---- multi-GPU plumbing without torch / NCCL: symmetric buffers, multicast mapping, device barrier -------------- */
/* One communicator per rank (a process, or a thread driving its own context) of ONE NVLink domain, world <= 8.  The only
 * thing the caller supplies is an all-gather of a few bytes for the bootstrap -- MPI_Allgather in a TensorToolkit program
 * (the reference distributes its DMRG mat-vec over MPI ranks, dmrg/contract_1sector.h:181-228), torch.distributed in the
 * Python harness, a std::barrier in a threaded test: gather `bytes` from every rank into recv[world * bytes], return 0. */
typedef int (*qlb200_allgather_fn)(void *user, const void *send, void *recv, size_t bytes);
typedef struct qlb200_comm qlb200_comm;
int qlb200_comm_create(qlb200_ctx *ctx, int32_t world, int32_t rank, qlb200_allgather_fn allgather, void *user, qlb200_comm **out);
void qlb200_comm_destroy(qlb200_comm *c);      /* collective */
int qlb200_comm_has_multicast(const qlb200_comm *c);
/* Collective: a buffer of `bytes` on every rank (CUDA virtual-memory management; handles travel as file descriptors over
 * abstract unix sockets).  local = this rank's buffer; peers[world] = every rank's buffer mapped into this process
 * (unicast loads / stores over NVLink; peers[rank] == local), for qlb200_execute_bcast / qlb200_fanout_copy; multicast, if
 * not NULL on entry, receives the NVSwitch multicast mapping of all of them (NULL when the fabric has none), for
 * qlb200_execute_mcast.  Zero-filled. */
int qlb200_comm_alloc(qlb200_comm *c, size_t bytes, void **local, void **peers, void **multicast);
int qlb200_comm_free(qlb200_comm *c, void *local);     /* collective */
/* Device-side barrier on the context's stream (a one-block kernel exchanging epochs through peer memory with system-scope
 * release / acquire): everything every rank enqueued before it -- e.g. the peer stores of a GEMM epilogue -- is visible to
 * everything any rank enqueues after it.  No host synchronisation; legal under stream capture. */
int qlb200_comm_barrier(qlb200_comm *c);

/* ---- CUDA graphs: replay a fixed sequence of executes (one Lanczos mat-vec) with one launch ---- */
/* Everything enqueued on the context's stream between begin and end -- this library's kernels and foreign work
 * such as an NCCL collective or a symmetric-memory barrier -- is captured (relaxed mode) instead of executed;
 * qlb200_graph_launch replays it.  Device-pointer executes only; run the sequence once before capturing so that
 * the workspace arena has its final size (an execute that would have to grow it during capture fails with
 * QLB200_ERR_UNSUPPORTED).  A graph keeps the arena addresses it was captured with: while any graph of a context is
 * alive, arenas outgrown by later plans are kept allocated (retired) instead of freed, so replay stays valid.  Destroy a
 * context's graphs before the context. */
typedef struct qlb200_graph qlb200_graph;
int qlb200_graph_begin(qlb200_ctx *ctx);
int qlb200_graph_end(qlb200_ctx *ctx, qlb200_graph **out);
int qlb200_graph_launch(qlb200_ctx *ctx, qlb200_graph *g);
void qlb200_graph_destroy(qlb200_graph *g);

/* number of kernel launches the last execute on this ctx issued */
uint64_t qlb200_ctx_launch_count(const qlb200_ctx *ctx);

/* ---- whole-tensor transpose (BlockSparseDataTensor::Transpose) ----------------------------- */
/* out block list: transposed shell (ascending new blk_idx), offsets; scale = fermion reorder sign. */
int qlb200_tplan_create(qlb200_ctx *ctx, const qlb200_shell *t, const int32_t *perm, int dtype,
                        qlb200_tplan **out);
void qlb200_tplan_destroy(qlb200_tplan *p);
uint64_t qlb200_tplan_nblk(const qlb200_tplan *p);
int qlb200_tplan_blocks(const qlb200_tplan *p, uint64_t *blk_idx, uint32_t *blk_coors,
                        uint32_t *shape, uint64_t *offset, int8_t *scale);
int qlb200_transpose_execute(qlb200_ctx *ctx, qlb200_tplan *p, const void *src, void *dst,
                             int mem_kind);

/* ---- matrix-free axis operations (dmrg/axis_ops.h) -------------------------------------------- */
/* out = in with one / two axes multiplied by rank-2 operators, AXIS ORDER PRESERVED:
 *     out[.., j1, .., j2, ..] = sum_{i1, i2} in[.., i1, .., i2, ..] * op1[i1, j1] * op2[i2, j2]
 * Replaces qlten::dmrg::ApplyRank2ToAxisPreserveOrder (tensor_manipulation/dmrg/axis_ops.h:2889-2992; per-block kernel
 * AddRank2AxisBlock :1695-1775) and ApplyTwoRank2ToAxesPreserveOrder (:2994-3125) -- bosonic quantum numbers only, as in
 * the reference.  op shells are rank 2 in the reference's {input_index, output_index} layout, op.index(0) ==
 * InverseIndex(in.index(axis)).  One kernel launch applies the operator(s) to every block in a single pass over the input
 * (no single-axis intermediate, no transposes).  The launch handles operator blocks up to 8 x 8 (site operators of a DMRG
 * Hamiltonian); larger ones return QLB200_ERR_UNSUPPORTED from qlb200_axis_plan_create and go through
 * qlb200_match_create + qlb200_tplan_create (contract, then move the new index into place) in the host adapters. */
typedef struct qlb200_axis qlb200_axis;           /* block pairing + output topology (host only) */
typedef struct qlb200_axis_plan qlb200_axis_plan; /* device tables */
int qlb200_axis_create(const qlb200_shell *in, int32_t nops, const qlb200_shell *op1, int32_t axis1, const qlb200_shell *op2,
                       int32_t axis2, qlb200_axis **out);
void qlb200_axis_destroy(qlb200_axis *a);
uint64_t qlb200_axis_out_nblk(const qlb200_axis *a);
uint64_t qlb200_axis_out_elems(const qlb200_axis *a);
uint64_t qlb200_axis_nterm(const qlb200_axis *a);            /* (input block, op block[s]) triples */
int qlb200_axis_out_blocks(const qlb200_axis *a, uint64_t *blk_idx, uint32_t *blk_coors, uint32_t *shape, uint64_t *offset);
int qlb200_axis_plan_create(qlb200_ctx *ctx, const qlb200_axis *a, int dtype, qlb200_axis_plan **out);
void qlb200_axis_plan_destroy(qlb200_axis_plan *p);
/* algorithmic traffic of one execute: every contributing input block once per output block it feeds, the output once */
int qlb200_axis_plan_bytes(const qlb200_axis_plan *p, uint64_t *read_bytes, uint64_t *write_bytes);
int qlb200_axis_execute(qlb200_ctx *ctx, qlb200_axis_plan *p, const void *in, const void *op1, const void *op2, void *out,
                        int mem_kind);

/* ---- batched range copy (multi-GPU: unpack all-gathered output slabs into the full layout) --- */
/* dst[dst_off[i] .. +len[i]) = src[src_off[i] .. +len[i]) for every i, one launch of the permute
 * kernel in its contiguous-run mode.  Offsets / lengths in elements; device pointers only. */
int qlb200_cplan_create(qlb200_ctx *ctx, int dtype, uint64_t n, const uint64_t *src_off,
                        const uint64_t *dst_off, const uint64_t *len, qlb200_tplan **out);
int qlb200_copy_execute(qlb200_ctx *ctx, qlb200_tplan *p, const void *src, void *dst);

#ifdef __cplusplus
}
#endif
#endif /* QLB200_H */
