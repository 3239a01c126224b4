"""oracle/refbridge.py -- TEST INFRASTRUCTURE ONLY.

ctypes bridge to oracle/_ref/libqlref.so: the UNMODIFIED reference (TensorToolkit CPU path, HPTT +
OpenBLAS) compiled by oracle/Makefile.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py may import this module; the product package never does.
"""
import ctypes as C
import os

import numpy as np

from tensortoolkit_b200.tensor import BlockSparseTensor, Index, KIND_ORDINAL

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libqlref.so")          # the unmodified reference, nothing else
ADAPTER_LIB = os.path.join(_HERE, "_ref", "libqladapter.so")  # include/qlten_b200/*.h instantiated on reference tensors (links libqlb200.so)

_I64P = C.POINTER(C.c_int64)
_P = C.c_void_p


def available() -> bool:
    return os.path.exists(REF_LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")   # see BASELINE.md: active spin-wait oversubscribes
        L = C.CDLL(REF_LIB, mode=C.RTLD_GLOBAL)
        sig = {
            "qlref_set_seed": (None, [C.c_uint64]),
            "qlref_set_threads": (None, [C.c_int]),
            "qlref_index_new": (_P, [C.c_int, C.c_int, C.c_int, _I64P, _I64P]),
            "qlref_index_free": (None, [_P]),
            "qlref_tensor_new": (_P, [C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
            "qlref_tensor_free": (None, [_P]),
            "qlref_tensor_clone": (_P, [_P]),
            "qlref_tensor_rank": (C.c_int, [_P]),
            "qlref_tensor_dtype": (C.c_int, [_P]),
            "qlref_tensor_is_default": (C.c_int, [_P]),
            "qlref_tensor_fermionic": (C.c_int, [_P]),
            "qlref_tensor_nblk": (C.c_uint64, [_P]),
            "qlref_tensor_raw_size": (C.c_uint64, [_P]),
            "qlref_tensor_raw": (_P, [_P]),
            "qlref_tensor_shell": (None, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_int8)]),
            "qlref_tensor_sectors": (None, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]),
            "qlref_tensor_blocks": (None, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
            "qlref_tensor_random": (None, [_P, _I64P]),
            "qlref_tensor_transpose": (None, [_P, _I64P]),
            "qlref_tensor_norm2": (C.c_double, [_P]),
            "qlref_tensor_indexes_equal": (C.c_int, [_P, _P]),
            "qlref_contract": (_P, [_P, _P, C.c_int, _I64P, _I64P]),
            "qlref_contract_time": (C.c_double, [_P, _P, C.c_int, _I64P, _I64P, C.c_int]),
            "qlref_contract_1sector": (_P, [_P, C.c_int64, C.c_int64, _P, C.c_int, _I64P, _I64P]),
            "qlref_contract_tasks": (C.c_uint64, [_P, _P, C.c_int, _I64P, _I64P, C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
            "qlref_contract_cost": (None, [_P, _P, C.c_int, _I64P, _I64P, C.POINTER(C.c_double)]),
            "qlref_tensor_write": (C.c_int, [_P, C.c_char_p]),
            "qlref_tensor_read": (C.c_int, [_P, C.c_char_p]),
            "qlref_contract_contiguous": (_P, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int]),
            "qlref_tensor_new_default": (_P, [_P]),
            "qlref_contract_accumulate": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double), _P,
                                                    C.c_int, C.POINTER(C.c_uint64)]),
            "qlref_apply_rank2": (_P, [_P, _P, C.c_int64, _P, C.c_int64]),
            "qlref_tensor_fill": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_uint32), _P, C.c_uint64]),
            "qlref_raw_contract": (C.c_double, [C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                                C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                                C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_double), _P, _P, _P, _P, _P]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


_adapter = None


def adapter():
    """The drop-in adapter library (qlref_b200_*): loaded only by tests that call the CUDA path through the C++ adapter,
    never by the reference arm / cpu_baseline leg of bench.py."""
    global _adapter
    if _adapter is None:
        lib()                                   # reference symbols first (the adapter library links against them)
        L = C.CDLL(ADAPTER_LIB, mode=C.RTLD_GLOBAL)
        sig = {
            "qlref_b200_contract": (_P, [_P, _P, C.c_int, _I64P, _I64P, _P]),
            "qlref_b200_contract_1sector": (_P, [_P, C.c_int64, C.c_int64, _P, C.c_int, _I64P, _I64P, _P]),
            "qlref_b200_transpose": (C.c_int, [_P, _I64P, _P]),
            "qlref_b200_contract_contiguous": (_P, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int, _P]),
            "qlref_b200_row_slab": (_P, [_P, C.c_int64, C.c_int64, _I64P]),
            "qlref_b200_cut_rows": (C.c_int, [_P, _P, C.c_int, _I64P, _I64P, C.c_int64, C.c_int, C.c_int, _I64P]),
            "qlref_b200_apply_rank2": (_P, [_P, _P, C.c_int64, _P, C.c_int64, _P]),
            "qlref_b200_contract_accumulate": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double), _P,
                                                         C.c_int, C.POINTER(C.c_uint64), _P]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _adapter = L
    return _adapter


def _i64(v):
    return (C.c_int64 * max(1, len(v)))(*[int(x) for x in v])


def set_seed(seed: int):
    lib().qlref_set_seed(seed)


def set_threads(n: int):
    lib().qlref_set_threads(n)


class RefTensor:
    """A QLTensor<double|complex, QNT> living inside the reference library."""

    def __init__(self, handle, indexes, dtype):
        self.h = handle
        self.indexes = list(indexes) if indexes is not None else None
        self.dtype = np.dtype(dtype)

    @staticmethod
    def new(indexes, dtype=np.float64, kind=None) -> "RefTensor":
        """`kind` names the quantum-number type of a rank-0 tensor (no index to read it from)."""
        L = lib()
        kind = indexes[0].kind if len(indexes) else kind
        ko = KIND_ORDINAL[kind.name]
        hs = []
        for ix in indexes:
            qn = [v for s in ix.sectors for v in s.qn]
            dg = [s.dgnc for s in ix.sectors]
            hs.append(L.qlref_index_new(ko, ix.dir, ix.nsct, _i64(qn), _i64(dg)))
        arr = (_P * len(hs))(*hs)
        t = L.qlref_tensor_new(ko, 0 if np.dtype(dtype) == np.float64 else 1, len(hs), arr)
        for h in hs:
            L.qlref_index_free(h)
        return RefTensor(t, indexes, dtype)

    def random(self, div):
        lib().qlref_tensor_random(self.h, _i64(list(div)))
        return self

    @staticmethod
    def from_bst(t: BlockSparseTensor, kind=None) -> "RefTensor":
        """A reference tensor with the same indexes, stored blocks and raw data as the product's host mirror."""
        r = RefTensor.new(t.indexes, t.dtype, kind)
        coors = np.ascontiguousarray(t.blk_coors, dtype=np.uint32)
        data = np.ascontiguousarray(t.data)
        if data.size:
            rc = lib().qlref_tensor_fill(r.h, t.nblk, coors.ctypes.data_as(C.POINTER(C.c_uint32)), data.ctypes.data, data.size)
            if rc != 0:
                raise RuntimeError("qlref_tensor_fill: block list and raw size disagree")
        return r

    def clone(self) -> "RefTensor":
        return RefTensor(lib().qlref_tensor_clone(self.h), self.indexes, self.dtype)

    def write_file(self, path: str):
        """`ofstream << tensor`: the reference's stream format (qltensor_impl.h:823-833)."""
        if lib().qlref_tensor_write(self.h, str(path).encode()) != 0:
            raise RuntimeError("reference could not write " + str(path))

    def read_file(self, path: str) -> "RefTensor":
        """`ifstream >> tensor` into this handle (same QN kind / element type); the caller's index list is kept as the
        expectation to compare against."""
        if lib().qlref_tensor_read(self.h, str(path).encode()) != 0:
            raise RuntimeError("reference could not read " + str(path))
        return self

    def indexes_equal(self, other: "RefTensor") -> bool:
        return bool(lib().qlref_tensor_indexes_equal(self.h, other.h))

    def transpose(self, perm):
        lib().qlref_tensor_transpose(self.h, _i64(perm))
        self.indexes = [self.indexes[i] for i in perm]
        return self

    def b200_transpose(self, perm, ctx_handle=None):
        rc = adapter().qlref_b200_transpose(self.h, _i64(perm), ctx_handle)
        if rc != 0:
            raise RuntimeError("qlten::b200::Transpose failed")
        self.indexes = [self.indexes[i] for i in perm]
        return self

    @property
    def rank(self):
        return lib().qlref_tensor_rank(self.h)

    @property
    def nblk(self):
        return int(lib().qlref_tensor_nblk(self.h))

    def is_default(self):
        return bool(lib().qlref_tensor_is_default(self.h))

    def norm2(self):
        return lib().qlref_tensor_norm2(self.h)

    def raw(self) -> np.ndarray:
        n = int(lib().qlref_tensor_raw_size(self.h))
        if n == 0:
            return np.zeros(0, self.dtype)
        p = lib().qlref_tensor_raw(self.h)
        buf = (C.c_char * (n * self.dtype.itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=self.dtype, count=n).copy()

    def blocks(self):
        n, r = self.nblk, self.rank
        idx = np.zeros(n, np.uint64); off = np.zeros(n, np.uint64)
        coors = np.zeros((n, r), np.uint32); shape = np.zeros((n, r), np.uint32)
        if n:
            lib().qlref_tensor_blocks(self.h, idx.ctypes.data_as(C.POINTER(C.c_uint64)), coors.ctypes.data_as(C.POINTER(C.c_uint32)),
                                      shape.ctypes.data_as(C.POINTER(C.c_uint32)), off.ctypes.data_as(C.POINTER(C.c_uint64)))
        return idx, coors, shape, off

    def shell_arrays(self):
        r = self.rank
        nsct = np.zeros(r, np.uint32); dirs = np.zeros(r, np.int8)
        lib().qlref_tensor_shell(self.h, nsct.ctypes.data_as(C.POINTER(C.c_uint32)), dirs.ctypes.data_as(C.POINTER(C.c_int8)))
        tot = int(nsct.sum())
        deg = np.zeros(tot, np.uint32); par = np.zeros(tot, np.uint8)
        lib().qlref_tensor_sectors(self.h, deg.ctypes.data_as(C.POINTER(C.c_uint32)), par.ctypes.data_as(C.POINTER(C.c_uint8)))
        return nsct, dirs, deg, par

    def to_bst(self) -> BlockSparseTensor:
        """Copy into the product's host mirror (same indexes, same block order, same raw data)."""
        t = BlockSparseTensor(self.indexes, self.dtype)
        if self.rank == 0:
            t.data = self.raw()
            return t
        _, coors, _, _ = self.blocks()
        t.set_blocks(coors, self.raw())
        return t

    def free(self):
        if self.h:
            lib().qlref_tensor_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _c_indexes(a: RefTensor, b: RefTensor, axes):
    sa = [i for i in range(len(a.indexes)) if i not in axes[0]]
    sb = [i for i in range(len(b.indexes)) if i not in axes[1]]
    return [a.indexes[i] for i in sa] + [b.indexes[i] for i in sb]


def contract(a: RefTensor, b: RefTensor, axes) -> RefTensor:
    h = lib().qlref_contract(a.h, b.h, len(axes[0]), _i64(axes[0]), _i64(axes[1]))
    return RefTensor(h, _c_indexes(a, b, axes), a.dtype)


def contract_time(a: RefTensor, b: RefTensor, axes, reps=1) -> float:
    return lib().qlref_contract_time(a.h, b.h, len(axes[0]), _i64(axes[0]), _i64(axes[1]), reps)


def contract_1sector(a: RefTensor, axis, sct, b: RefTensor, axes) -> RefTensor:
    h = lib().qlref_contract_1sector(a.h, axis, sct, b.h, len(axes[0]), _i64(axes[0]), _i64(axes[1]))
    return RefTensor(h, _c_indexes(a, b, axes), a.dtype)


def b200_contract(a: RefTensor, b: RefTensor, axes, ctx_handle=None) -> RefTensor:
    """qlten::b200::Contract on reference QLTensors (the drop-in adapter, CUDA path)."""
    h = adapter().qlref_b200_contract(a.h, b.h, len(axes[0]), _i64(axes[0]), _i64(axes[1]), ctx_handle)
    if not h:
        raise RuntimeError("qlten::b200::Contract failed")
    return RefTensor(h, _c_indexes(a, b, axes), a.dtype)


def b200_row_slab(t: RefTensor, axis: int, ranges, new_indexes) -> RefTensor:
    """qlten::b200::RowSlab on a reference QLTensor (host only); `new_indexes` = the index list the caller expects."""
    flat = [int(x) for r in ranges for x in r]
    h = adapter().qlref_b200_row_slab(t.h, axis, len(ranges), _i64(flat))
    if not h:
        raise RuntimeError("qlten::b200::RowSlab failed")
    return RefTensor(h, new_indexes, t.dtype)


def b200_cut_rows(a: RefTensor, b: RefTensor, axes, split_axis: int, world: int, snap: int = 8):
    """qlten::b200::SectorFlops + CutRowLine for one contraction: [rank][sector] -> (lo, hi)."""
    nsct = a.indexes[split_axis].nsct
    out = (C.c_int64 * (world * nsct * 2))()
    rc = adapter().qlref_b200_cut_rows(a.h, b.h, len(axes[0]), _i64(axes[0]), _i64(axes[1]), split_axis, world, snap, out)
    if rc != 0:
        raise RuntimeError("qlten::b200::CutRowLine failed")
    return [[(int(out[(r * nsct + s) * 2]), int(out[(r * nsct + s) * 2 + 1])) for s in range(nsct)] for r in range(world)]


def b200_contract_1sector(a: RefTensor, axis, sct, b: RefTensor, axes, ctx_handle=None) -> RefTensor:
    h = adapter().qlref_b200_contract_1sector(a.h, axis, sct, b.h, len(axes[0]), _i64(axes[0]), _i64(axes[1]), ctx_handle)
    if not h:
        raise RuntimeError("qlten::b200::Contract1Sector failed")
    return RefTensor(h, _c_indexes(a, b, axes), a.dtype)


def _c_indexes_cyclic(a: RefTensor, b: RefTensor, a_start, b_start, size):
    ra, rb = len(a.indexes), len(b.indexes)
    sa = [(a_start + size + i) % ra for i in range(ra - size)]
    sb = [(b_start + size + i) % rb for i in range(rb - size)]
    return [a.indexes[i] for i in sa] + [b.indexes[i] for i in sb]


SIDES = {("tail", "head"): 0, ("head", "head"): 1, ("tail", "tail"): 2, ("head", "tail"): 3}


def contract_contiguous(a: RefTensor, b: RefTensor, a_start: int, b_start: int, size: int, sides=("tail", "head")) -> RefTensor:
    """The reference's qlten::ContractContiguousAxes<ASide, BSide> (contract_contiguous_axes.h:849-873)."""
    L = lib()
    h = L.qlref_contract_contiguous(a.h, b.h, int(a_start), int(b_start), int(size), SIDES[tuple(sides)])
    return RefTensor(h, _c_indexes_cyclic(a, b, a_start, b_start, size), a.dtype)


def b200_contract_contiguous(a: RefTensor, b: RefTensor, a_start: int, b_start: int, size: int, sides=("tail", "head"),
                             ctx_handle=None) -> RefTensor:
    """qlten::b200::ContractContiguousAxes on reference tensors (the drop-in adapter over the C ABI)."""
    L = adapter()
    h = L.qlref_b200_contract_contiguous(a.h, b.h, int(a_start), int(b_start), int(size), SIDES[tuple(sides)], ctx_handle)
    if not h:
        raise RuntimeError("qlref_b200_contract_contiguous failed (see stderr)")
    return RefTensor(h, _c_indexes_cyclic(a, b, a_start, b_start, size), a.dtype)


ACCUM_STAT_NAMES = ("raw_data_contract_tasks", "gemm_calls", "accumulate_calls", "accumulate_gemm_calls", "output_tensor_rebuilds",
                    "temporary_output_bytes_avoided", "output_topology_expansions", "output_expand_copy_bytes", "output_expand_new_blocks",
                    "output_untouched_scale_bytes")


def default_like(t: RefTensor) -> RefTensor:
    """A default-constructed QLTensor of t's element / quantum-number type."""
    return RefTensor(lib().qlref_tensor_new_default(t.h), None, t.dtype)


def _accumulate(fn, a, b, a_start, b_start, size, alpha, beta, c, try_only, extra=()):
    """Shared driver of the reference's / the adapter's ContractTailHeadContiguousAccumulate on handle `c` (updated in place).
    Returns (ok, stats dict); raises RuntimeError when the callee threw."""
    al = (C.c_double * 2)(complex(alpha).real, complex(alpha).imag)
    be = (C.c_double * 2)(complex(beta).real, complex(beta).imag)
    st = (C.c_uint64 * 10)()
    rc = fn(a.h, b.h, int(a_start), int(b_start), int(size), al, be, c.h, int(try_only), st, *extra)
    if rc < 0:
        raise RuntimeError("ContractTailHeadContiguousAccumulate threw (see stderr)")
    c.indexes = _c_indexes_cyclic(a, b, a_start, b_start, size)
    return bool(rc), dict(zip(ACCUM_STAT_NAMES, (int(x) for x in st)))


def contract_accumulate(a, b, a_start, b_start, size, alpha, beta, c: RefTensor, try_only=False):
    """The reference's (Try)ContractTailHeadContiguousAccumulate (contract_contiguous_axes.h:954-1041) on handle c."""
    return _accumulate(lib().qlref_contract_accumulate, a, b, a_start, b_start, size, alpha, beta, c, try_only)


def b200_contract_accumulate(a, b, a_start, b_start, size, alpha, beta, c: RefTensor, try_only=False, ctx_handle=None):
    """qlten::b200::(Try)ContractTailHeadContiguousAccumulate on reference tensors (the drop-in adapter)."""
    return _accumulate(adapter().qlref_b200_contract_accumulate, a, b, a_start, b_start, size, alpha, beta, c, try_only, (ctx_handle,))


def _axis_indexes(x, ops):
    idxs = list(x.indexes)
    for op, ax in ops:
        idxs[ax] = op.indexes[1]
    return idxs


def apply_rank2(x: RefTensor, ops) -> RefTensor:
    """The reference's dmrg::ApplyRank2ToAxisPreserveOrder (one (op, axis)) / ApplyTwoRank2ToAxesPreserveOrder (two)."""
    (o1, a1), (o2, a2) = ops[0], (ops[1] if len(ops) > 1 else (None, 0))
    h = lib().qlref_apply_rank2(x.h, o1.h, int(a1), o2.h if o2 is not None else None, int(a2))
    if not h:
        raise RuntimeError("reference axis operation failed (see stderr)")
    return RefTensor(h, _axis_indexes(x, ops), x.dtype)


def b200_apply_rank2(x: RefTensor, ops, ctx_handle=None) -> RefTensor:
    """qlten::b200::dmrg::Apply(Two)Rank2To...PreserveOrder on reference tensors (the drop-in adapter)."""
    (o1, a1), (o2, a2) = ops[0], (ops[1] if len(ops) > 1 else (None, 0))
    h = adapter().qlref_b200_apply_rank2(x.h, o1.h, int(a1), o2.h if o2 is not None else None, int(a2), ctx_handle)
    if not h:
        raise RuntimeError("qlten::b200::dmrg axis operation failed (see stderr)")
    return RefTensor(h, _axis_indexes(x, ops), x.dtype)


def raw_contract(dtype, a_rank, a_perm, a_shape, a_off, b_rank, b_perm, b_shape, b_off, tasks, A, B, c_elems, keep_permuted=False):
    """The reference's executor loop (CtrctTwoBSDTAndAssignIn, global_operations.h:895-992) on bare descriptor tables:
    hp_numeric::TensorTranspose per distinct block + hp_numeric::MatMultiply per task.  `tasks`: structured array /
    sequence with a_ord, b_ord, a_off, b_off, c_off, m, k, n, sign, first.  a_perm / b_perm None = no transpose.
    Returns (C, seconds[, A_permuted, B_permuted]) -- permuted block b lies at a_off[b] of A_permuted."""
    dtype = np.dtype(dtype)
    n = len(tasks)
    tu = np.zeros((n, 8), np.uint64); ts = np.zeros((n, 2), np.float64)
    for i, t in enumerate(tasks):
        tu[i] = (t["a_ord"], t["b_ord"], t["a_off"], t["b_off"], t["c_off"], t["m"], t["k"], t["n"])
        ts[i] = (t["sign"], 0.0 if t["first"] else 1.0)
    ash = np.ascontiguousarray(np.asarray(a_shape, np.uint32).reshape(-1)); bsh = np.ascontiguousarray(np.asarray(b_shape, np.uint32).reshape(-1))
    aof = np.ascontiguousarray(np.asarray(a_off, np.uint64)); bof = np.ascontiguousarray(np.asarray(b_off, np.uint64))
    A = np.ascontiguousarray(A, dtype=dtype); B = np.ascontiguousarray(B, dtype=dtype)
    Cout = np.zeros(c_elems, dtype)
    At = np.zeros_like(A) if keep_permuted and a_perm is not None else None
    Bt = np.zeros_like(B) if keep_permuted and b_perm is not None else None
    i32 = lambda v: (C.c_int32 * len(v))(*[int(x) for x in v]) if v is not None else None
    u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    sec = lib().qlref_raw_contract(0 if dtype == np.float64 else 1, a_rank, i32(a_perm), ash.ctypes.data_as(u32p), aof.ctypes.data_as(u64p),
                                   b_rank, i32(b_perm), bsh.ctypes.data_as(u32p), bof.ctypes.data_as(u64p), n,
                                   tu.ctypes.data_as(u64p), ts.ctypes.data_as(C.POINTER(C.c_double)), A.ctypes.data, B.ctypes.data,
                                   Cout.ctypes.data, At.ctypes.data if At is not None else None, Bt.ctypes.data if Bt is not None else None)
    return (Cout, sec, At, Bt) if keep_permuted else (Cout, sec)


def contract_tasks(a: RefTensor, b: RefTensor, axes, sorted_by_c=False):
    """The reference's RawDataCtrctTask list as (uint64[n,9], float64[n,2]) =
    (a_idx,b_idx,c_idx,a_off,b_off,c_off,m,k,n), (f_ex_sign, beta)."""
    L = lib()
    n = int(L.qlref_contract_tasks(a.h, b.h, len(axes[0]), _i64(axes[0]), _i64(axes[1]), int(sorted_by_c), 0, None, None))
    u = np.zeros((n, 9), np.uint64); d = np.zeros((n, 2), np.float64)
    if n:
        L.qlref_contract_tasks(a.h, b.h, len(axes[0]), _i64(axes[0]), _i64(axes[1]), int(sorted_by_c), n,
                               u.ctypes.data_as(C.POINTER(C.c_uint64)), d.ctypes.data_as(C.POINTER(C.c_double)))
    return u, d


def contract_cost(a: RefTensor, b: RefTensor, axes) -> dict:
    out = (C.c_double * 8)()
    lib().qlref_contract_cost(a.h, b.h, len(axes[0]), _i64(axes[0]), _i64(axes[1]), out)
    keys = ["flops", "gemm_count", "candidate_block_pair_count", "output_block_count", "output_raw_elem_count",
            "read_bytes", "write_bytes", "temp_peak_bytes"]
    return dict(zip(keys, list(out)))
