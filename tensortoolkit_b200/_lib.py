"""ctypes binding of the C ABI in include/qlb200.h (tensortoolkit_b200/libqlb200.so).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``make -C tensortoolkit_b200/csrc``.
There is no CPU fallback: if the library is missing, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# QLB200_LIB: load another build of the same library (kernel A/B experiments, exp/); default = the in-tree build
LIB_PATH = os.environ.get("QLB200_LIB") or os.path.join(_HERE, "libqlb200.so")

QLB200_MAX_RANK = 8
OK = 0
ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED, ERR_LAYOUT = -1, -2, -3, -4, -5
F64, C64 = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
SPLIT_BY_A, SPLIT_BY_B, SPLIT_BY_C = 0, 1, 2
DIR_IN, DIR_OUT = -1, 1
PLAN_DETERMINISTIC, PLAN_NO_SKINNY, PLAN_PERMUTE_ALL, PLAN_NO_SPLIT_K, PLAN_CPLX_4M, PLAN_STAGGER_OUTPUT, PLAN_STREAM_K = 1, 2, 8, 16, 32, 64, 128
PLAN_NO_VIEW = 256


class Shell(C.Structure):
    _fields_ = [
        ("rank", C.c_int32),
        ("nsct", C.POINTER(C.c_uint32)),
        ("deg", C.POINTER(C.c_uint32)),
        ("parity", C.POINTER(C.c_uint8)),
        ("dir", C.POINTER(C.c_int8)),
        ("nblk", C.c_uint64),
        ("blk_coors", C.POINTER(C.c_uint32)),
    ]


class Task(C.Structure):
    _fields_ = [
        ("a_blk_idx", C.c_uint64), ("b_blk_idx", C.c_uint64), ("c_blk_idx", C.c_uint64),
        ("a_off", C.c_uint64), ("b_off", C.c_uint64), ("c_off", C.c_uint64),
        ("a_ord", C.c_uint32), ("b_ord", C.c_uint32), ("c_ord", C.c_uint32),
        ("m", C.c_uint32), ("k", C.c_uint32), ("n", C.c_uint32),
        ("sign", C.c_int8), ("first", C.c_uint8), ("pad_", C.c_uint8 * 2),
    ]


class Cost(C.Structure):
    _fields_ = [
        ("flops", C.c_double), ("gemm_count", C.c_uint64), ("candidate_block_pair_count", C.c_uint64),
        ("output_block_count", C.c_uint64), ("output_raw_elem_count", C.c_uint64),
        ("read_bytes", C.c_uint64), ("write_bytes", C.c_uint64), ("temp_peak_bytes", C.c_uint64),
    ]


class PlanStats(C.Structure):
    _fields_ = [
        ("flops", C.c_double), ("ntask", C.c_uint64), ("ngroup", C.c_uint64), ("ntile_dmma", C.c_uint64),
        ("nrow_skinny", C.c_uint64), ("permute_elems_a", C.c_uint64), ("permute_elems_b", C.c_uint64),
        ("workspace_bytes", C.c_uint64), ("gemm_read_bytes", C.c_uint64), ("gemm_write_bytes", C.c_uint64),
    ]


class AccumStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "raw_data_contract_tasks", "gemm_calls", "accumulate_calls", "accumulate_gemm_calls", "output_tensor_rebuilds",
        "temporary_output_bytes_avoided", "output_topology_expansions", "output_expand_copy_bytes", "output_expand_new_blocks",
        "output_untouched_scale_bytes")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)


class Piece(C.Structure):
    _fields_ = [("sector", C.c_uint32), ("lo", C.c_uint32), ("hi", C.c_uint32), ("pad_", C.c_uint32), ("weight", C.c_double)]


class SlabInfo(C.Structure):
    _fields_ = [("nsct_kept", C.c_uint32), ("pad_", C.c_uint32), ("nblk_kept", C.c_uint64), ("elems", C.c_uint64), ("ncopy", C.c_uint64)]


class Item(C.Structure):
    _fields_ = [("group", C.c_uint32), ("row0", C.c_uint32), ("rows", C.c_uint32), ("n", C.c_uint32)]


class Unit(C.Structure):
    _fields_ = [("group", C.c_uint32), ("tm", C.c_uint32), ("tn", C.c_uint32), ("s_begin", C.c_uint32), ("s_end", C.c_uint32),
                ("split", C.c_uint32), ("nsplit", C.c_uint32), ("rows", C.c_uint32), ("cols", C.c_uint32)]


# name -> (restype, argtypes); every symbol declared in include/qlb200.h
_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_I32P = C.POINTER(C.c_int32)
_U32P = C.POINTER(C.c_uint32)
_U64P = C.POINTER(C.c_uint64)
_I8P = C.POINTER(C.c_int8)
_SH = C.POINTER(Shell)
SYMBOLS = {
    "qlb200_version": (C.c_char_p, []),
    "qlb200_last_error": (C.c_char_p, []),
    "qlb200_match_create": (C.c_int, [_SH, _SH, C.c_int32, _I32P, _I32P, _PP]),
    "qlb200_match_destroy": (None, [_P]),
    "qlb200_match_create_1sector": (C.c_int, [_SH, C.c_int32, C.c_uint32, _SH, C.c_int32, _I32P, _I32P, _PP]),
    "qlb200_match_create_contiguous": (C.c_int, [_SH, _SH, C.c_int32, C.c_int32, C.c_int32, _PP]),
    "qlb200_match_saved_axes": (C.c_int32, [_P, C.c_int, _I32P]),
    "qlb200_match_c_rank": (C.c_int32, [_P]),
    "qlb200_match_c_nblk": (C.c_uint64, [_P]),
    "qlb200_match_c_elems": (C.c_uint64, [_P]),
    "qlb200_match_ntask": (C.c_uint64, [_P]),
    "qlb200_match_is_scalar": (C.c_int, [_P]),
    "qlb200_match_perm": (C.c_int, [_P, C.c_int, _I32P]),
    "qlb200_match_c_blocks": (C.c_int, [_P, _U64P, _U32P, _U32P, _U64P]),
    "qlb200_match_tasks": (C.c_int, [_P, C.c_int, C.POINTER(Task)]),
    "qlb200_estimate_cost": (C.c_int, [_P, C.c_int, C.POINTER(Cost)]),
    "qlb200_ctx_create": (C.c_int, [C.c_int, _PP]),
    "qlb200_ctx_destroy": (None, [_P]),
    "qlb200_ctx_sync": (C.c_int, [_P]),
    "qlb200_ctx_stream": (_P, [_P]),
    "qlb200_ctx_set_stream": (C.c_int, [_P, _P]),
    "qlb200_dev_alloc": (C.c_int, [_P, C.c_size_t, _PP]),
    "qlb200_dev_free": (C.c_int, [_P, _P]),
    "qlb200_memcpy_h2d": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "qlb200_memcpy_d2h": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "qlb200_host_register": (C.c_int, [_P, C.c_size_t]),
    "qlb200_host_unregister": (C.c_int, [_P]),
    "qlb200_plan_create": (C.c_int, [_P, _P, _SH, _SH, C.c_int, C.c_uint32, _PP]),
    "qlb200_plan_create_raw": (C.c_int, [_P, C.c_int, C.c_uint32, C.c_int32, _I32P, C.c_uint64, _U32P, _U64P,
                                         C.c_int32, _I32P, C.c_uint64, _U32P, _U64P, C.c_uint64, C.POINTER(Task),
                                         C.c_uint64, _PP]),
    "qlb200_plan_destroy": (None, [_P]),
    "qlb200_plan_partition": (C.c_int, [_P, C.c_int32, C.c_int32]),
    "qlb200_plan_c_range_count": (C.c_uint64, [_P]),
    "qlb200_plan_c_ranges": (C.c_int, [_P, _U64P, _U64P]),
    "qlb200_plan_get_stats": (C.c_int, [_P, C.POINTER(PlanStats)]),
    "qlb200_plan_operand_block": (C.c_int, [_P, C.c_int, C.c_uint64, _U64P]),
    "qlb200_plan_read_workspace": (C.c_int, [_P, _P, C.c_int, C.c_uint64, C.c_uint64, _P]),
    "qlb200_plan_units": (C.c_uint64, [_P, C.c_uint64, C.POINTER(Unit), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "qlb200_plan_items": (C.c_uint64, [_P, C.c_uint64, C.POINTER(Item)]),
    "qlb200_shard_sector_flops": (C.c_int, [_P, C.c_int32, C.c_int, C.POINTER(C.c_double)]),
    "qlb200_shard_restrict": (C.c_int, [_P, C.c_int32, _U32P, C.POINTER(SlabInfo), _U32P, _U32P, _U32P, _U32P, _U64P, _U64P, _U64P]),
    "qlb200_shard_cut_line": (C.c_int, [C.POINTER(Piece), C.c_uint64, _U32P, C.c_uint32, C.c_int32, C.c_int32, _U32P]),
    "qlb200_shard_reweigh": (C.c_uint64, [C.POINTER(Piece), C.c_uint64, _U32P, C.c_uint32, C.c_int32, C.POINTER(C.c_double), C.c_double,
                                          C.c_uint64, C.POINTER(Piece)]),
    "qlb200_plan_segments": (C.c_uint64, [_P, C.c_uint64, C.POINTER(C.c_uint32)]),
    "qlb200_execute": (C.c_int, [_P, _P, _P, _P, _P, C.c_int]),
    "qlb200_execute_permute": (C.c_int, [_P, _P, _P, _P]),
    "qlb200_execute_gemm": (C.c_int, [_P, _P, _P, _P, _P]),
    "qlb200_plan_split": (C.c_int, [_P, C.c_int, C.c_int32, C.POINTER(C.c_double), _PP, _U64P]),
    "qlb200_hostpipe_create": (C.c_int, [_P, _P, C.c_int, C.c_int32, C.POINTER(C.c_double), _P, C.c_int32, C.POINTER(C.c_double), _PP]),
    "qlb200_hostpipe_destroy": (None, [_P]),
    "qlb200_hostpipe_begin": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "qlb200_hostpipe_end": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "qlb200_hostpipe_launches": (C.c_uint64, [_P]),
    "qlb200_axis_create": (C.c_int, [_SH, C.c_int32, _SH, C.c_int32, _SH, C.c_int32, _PP]),
    "qlb200_axis_destroy": (None, [_P]),
    "qlb200_axis_out_nblk": (C.c_uint64, [_P]),
    "qlb200_axis_out_elems": (C.c_uint64, [_P]),
    "qlb200_axis_nterm": (C.c_uint64, [_P]),
    "qlb200_axis_out_blocks": (C.c_int, [_P, _U64P, _U32P, _U32P, _U64P]),
    "qlb200_axis_plan_create": (C.c_int, [_P, _P, C.c_int, _PP]),
    "qlb200_axis_plan_destroy": (None, [_P]),
    "qlb200_axis_plan_bytes": (C.c_int, [_P, _U64P, _U64P]),
    "qlb200_axis_execute": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int]),
    "qlb200_accum_create": (C.c_int, [_P, _SH, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), _PP]),
    "qlb200_accum_destroy": (None, [_P]),
    "qlb200_accum_nblk": (C.c_uint64, [_P]),
    "qlb200_accum_elems": (C.c_uint64, [_P]),
    "qlb200_accum_expanded": (C.c_int, [_P]),
    "qlb200_accum_blocks": (C.c_int, [_P, _U64P, _U32P, _U32P, _U64P, _U64P, C.POINTER(C.c_uint8)]),
    "qlb200_accum_get_stats": (C.c_int, [_P, C.POINTER(AccumStats)]),
    "qlb200_plan_create_accum": (C.c_int, [_P, _P, _P, C.c_int, C.c_uint32, _PP]),
    "qlb200_execute_accum": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int]),
    "qlb200_execute_bcast": (C.c_int, [_P, _P, _P, _P, _PP, C.c_int32]),
    "qlb200_execute_mcast": (C.c_int, [_P, _P, _P, _P, _P]),
    "qlb200_graph_begin": (C.c_int, [_P]),
    "qlb200_graph_end": (C.c_int, [_P, _PP]),
    "qlb200_graph_launch": (C.c_int, [_P, _P]),
    "qlb200_graph_destroy": (None, [_P]),
    "qlb200_comm_create": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, _PP]),
    "qlb200_comm_destroy": (None, [_P]),
    "qlb200_comm_has_multicast": (C.c_int, [_P]),
    "qlb200_comm_alloc": (C.c_int, [_P, C.c_size_t, _PP, _PP, _PP]),
    "qlb200_comm_free": (C.c_int, [_P, _P]),
    "qlb200_comm_barrier": (C.c_int, [_P]),
    "qlb200_fanout_copy": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _PP, C.c_int32, _P]),
    "qlb200_plan_remap_output": (C.c_int, [_P, C.c_uint64, _U64P, _U64P]),
    "qlb200_ipc_export": (C.c_int, [_P, _P, C.c_char_p]),
    "qlb200_ipc_open": (C.c_int, [_P, C.c_char_p, _PP]),
    "qlb200_ipc_close": (C.c_int, [_P, _P]),
    "qlb200_ctx_launch_count": (C.c_uint64, [_P]),
    "qlb200_tplan_create": (C.c_int, [_P, _SH, _I32P, C.c_int, _PP]),
    "qlb200_tplan_destroy": (None, [_P]),
    "qlb200_tplan_nblk": (C.c_uint64, [_P]),
    "qlb200_tplan_blocks": (C.c_int, [_P, _U64P, _U32P, _U32P, _U64P, _I8P]),
    "qlb200_transpose_execute": (C.c_int, [_P, _P, _P, _P, C.c_int]),
    "qlb200_cplan_create": (C.c_int, [_P, C.c_int, C.c_uint64, _U64P, _U64P, _U64P, _PP]),
    "qlb200_copy_execute": (C.c_int, [_P, _P, _P, _P]),
}


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the contraction path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


class QLB200Error(RuntimeError):
    pass


def check(rc, what):
    if rc != OK:
        raise QLB200Error(f"{what} failed ({rc}): {lib.qlb200_last_error().decode()}")
