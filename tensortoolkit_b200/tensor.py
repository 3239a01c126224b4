"""Host-side mirror of the reference's block-sparse tensor types, just enough to drive the
contraction path from Python (benchmarks, tests).  Names follow the reference:

  QNSector   include/qlten/qltensor/qnsct.h:35        (qn, degeneracy)
  Index      include/qlten/qltensor/index.h:46        (sector list + direction)
  BlockSparseTensor  ~ QLTensor + BlockSparseDataTensor
             include/qlten/qltensor/qltensor.h:61, blk_spar_data_ten/blk_spar_data_ten.h:53
             one flat raw buffer, blocks in ascending blk_idx order (row-major index of the block
             coordinates over the per-index sector counts), row-major inside a block.

Quantum numbers are tuples of ints under a named symmetry (``QNKind``); only what the matcher
needs is modelled: addition (for Div / Random), fermion parity, equality.
"""
import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib

IN, OUT = _lib.DIR_IN, _lib.DIR_OUT


@dataclass(frozen=True)
class QNKind:
    """Symmetry of a quantum number: special_qn types of the reference (qltensor/special_qn/)."""
    name: str
    nvals: int
    modulus: Tuple[int, ...]     # 0 = U(1) component, n = Z_n component
    fermionic: bool

    def parity(self, qn) -> int:
        # fU1QN: val % 2 (fu1qn.h:55); fU1U1QN: vals[0] % 2 (fu1u1qn.h:51); fZ2QN: val (fz2qn.h:52)
        return int(qn[0] % 2 != 0) if self.fermionic else 0

    def norm(self, qn):
        return tuple(int(v) % m if m else int(v) for v, m in zip(qn, self.modulus))


U1 = QNKind("U1QN", 1, (0,), False)
fU1 = QNKind("fU1QN", 1, (0,), True)
U1U1 = QNKind("U1U1QN", 2, (0, 0), False)
fU1U1 = QNKind("fU1U1QN", 2, (0, 0), True)
Z2 = QNKind("Z2QN", 1, (2,), False)
fZ2 = QNKind("fZ2QN", 1, (2,), True)
# ordinal used by the oracle bridge (oracle/ref_driver.cc: enum QNKind)
KIND_ORDINAL = {"U1QN": 0, "fU1QN": 1, "U1U1QN": 2, "fU1U1QN": 3, "Z2QN": 4, "fZ2QN": 5}


@dataclass(frozen=True)
class QNSector:
    qn: Tuple[int, ...]
    dgnc: int


class Index:
    def __init__(self, kind: QNKind, sectors: Sequence[QNSector], direction: int):
        assert direction in (IN, OUT)
        self.kind = kind
        self.sectors = tuple(QNSector(kind.norm(s.qn), int(s.dgnc)) for s in sectors)
        self.dir = direction

    @staticmethod
    def make(kind: QNKind, qn_dgnc: Sequence[Tuple[Sequence[int], int]], direction: int) -> "Index":
        return Index(kind, [QNSector(tuple(q) if not np.isscalar(q) else (int(q),), d) for q, d in qn_dgnc], direction)

    def inverse(self) -> "Index":
        """InverseIndex (index.h): same sectors, opposite direction."""
        return Index(self.kind, self.sectors, -self.dir)

    @property
    def nsct(self) -> int:
        return len(self.sectors)

    @property
    def dim(self) -> int:
        return sum(s.dgnc for s in self.sectors)

    def degs(self) -> np.ndarray:
        return np.array([s.dgnc for s in self.sectors], dtype=np.uint32)

    def __eq__(self, other):
        return isinstance(other, Index) and self.kind == other.kind and self.sectors == other.sectors and self.dir == other.dir

    def __hash__(self):
        return hash((self.kind.name, self.sectors, self.dir))


def _dtype_code(dtype) -> int:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return _lib.F64
    if dtype == np.complex128:
        return _lib.C64
    raise TypeError("only float64 / complex128 tensors are supported (reference ElemT double / complex<double>)")


class ShellView:
    """Keeps the numpy arrays a qlb200_shell points into alive."""

    def __init__(self, rank, nsct, deg, parity, dirs, coors):
        self.nsct = np.ascontiguousarray(nsct, dtype=np.uint32)
        self.deg = np.ascontiguousarray(deg, dtype=np.uint32)
        self.parity = None if parity is None else np.ascontiguousarray(parity, dtype=np.uint8)
        self.dirs = np.ascontiguousarray(dirs, dtype=np.int8)
        self.coors = np.ascontiguousarray(coors, dtype=np.uint32).reshape(-1)
        s = _lib.Shell()
        s.rank = rank
        s.nsct = self.nsct.ctypes.data_as(C.POINTER(C.c_uint32))
        s.deg = self.deg.ctypes.data_as(C.POINTER(C.c_uint32))
        s.parity = None if self.parity is None else self.parity.ctypes.data_as(C.POINTER(C.c_uint8))
        s.dir = self.dirs.ctypes.data_as(C.POINTER(C.c_int8))
        s.nblk = (len(self.coors) // rank) if rank else 0
        s.blk_coors = self.coors.ctypes.data_as(C.POINTER(C.c_uint32))
        self.struct = s

    def ptr(self):
        return C.byref(self.struct)


class BlockSparseTensor:
    """Block structure + one flat raw buffer (host numpy array)."""

    def __init__(self, indexes: Sequence[Index], dtype=np.float64):
        self.indexes: List[Index] = list(indexes)
        self.dtype = np.dtype(dtype)
        _dtype_code(self.dtype)
        self.blk_coors = np.zeros((0, self.rank), dtype=np.uint32)   # ascending blk_idx order
        self.data = np.zeros(0, dtype=self.dtype)
        self._refresh()

    # ---- structure ------------------------------------------------------------------------
    @property
    def rank(self) -> int:
        return len(self.indexes)

    @property
    def kind(self) -> QNKind:
        return self.indexes[0].kind

    @property
    def nblk(self) -> int:
        return self.blk_coors.shape[0]

    def nsct(self) -> np.ndarray:
        return np.array([ix.nsct for ix in self.indexes], dtype=np.uint32)

    def _refresh(self):
        r = self.rank
        if r == 0 or self.nblk == 0:
            self.blk_shape = np.zeros((0, r), dtype=np.uint32)
            self.blk_size = np.zeros(0, dtype=np.uint64)
            self.blk_offset = np.zeros(0, dtype=np.uint64)
            self.blk_idx = np.zeros(0, dtype=np.uint64)
            return
        degs = [ix.degs() for ix in self.indexes]
        self.blk_shape = np.stack([degs[i][self.blk_coors[:, i]] for i in range(r)], axis=1).astype(np.uint32)
        self.blk_size = np.prod(self.blk_shape.astype(np.uint64), axis=1)
        self.blk_offset = np.concatenate([[0], np.cumsum(self.blk_size)[:-1]]).astype(np.uint64)
        idx = np.zeros(self.nblk, dtype=np.uint64)
        for i in range(r):
            idx = idx * np.uint64(self.indexes[i].nsct) + self.blk_coors[:, i].astype(np.uint64)
        self.blk_idx = idx

    def set_blocks(self, coors: np.ndarray, data: np.ndarray = None):
        coors = np.asarray(coors, dtype=np.uint32).reshape(-1, self.rank)
        # sort into ascending blk_idx order
        if len(coors):
            order = np.lexsort(coors.T[::-1])
            coors = coors[order]
        self.blk_coors = coors
        self._refresh()
        n = int(self.blk_size.sum()) if self.nblk else 0
        if data is None:
            self.data = np.empty(n, dtype=self.dtype)
        else:
            assert data.size == n
            self.data = np.ascontiguousarray(data, dtype=self.dtype).reshape(-1)

    def div_blocks(self, div: Sequence[int]) -> np.ndarray:
        """All block coordinates whose quantum-number flow equals `div`
        (QLTensor::Random, qltensor_impl.h:375-403; CalcDiv, index.h:265-290)."""
        kind = self.kind
        div = np.array(kind.norm(tuple(div)), dtype=np.int64)
        total = None
        for ix in self.indexes:
            q = np.array([s.qn for s in ix.sectors], dtype=np.int64) * ix.dir     # [nsct, nv]
            total = q if total is None else (total[:, None, :] + q[None, :, :]).reshape(-1, kind.nvals)
        for v, m in enumerate(kind.modulus):
            if m:
                total[:, v] %= m
        hit = np.nonzero(np.all(total == div[None, :], axis=1))[0]
        return np.stack(np.unravel_index(hit, tuple(int(x) for x in self.nsct())), axis=1).astype(np.uint32)

    def random(self, div: Sequence[int], rng: np.random.Generator):
        """Fill with uniform [0,1) values in every block of divergence `div`."""
        self.set_blocks(self.div_blocks(div))
        n = self.data.size
        if self.dtype == np.complex128:
            self.data = (rng.random(n) + 1j * rng.random(n)).astype(np.complex128)
        else:
            self.data = rng.random(n)
        return self

    def shell(self) -> ShellView:
        kind = self.kind if self.rank else None
        deg = np.concatenate([ix.degs() for ix in self.indexes]) if self.rank else np.zeros(0, np.uint32)
        parity = None
        if kind is not None and kind.fermionic:
            parity = np.array([kind.parity(s.qn) for ix in self.indexes for s in ix.sectors], dtype=np.uint8)
        return ShellView(self.rank, self.nsct(), deg, parity, [ix.dir for ix in self.indexes], self.blk_coors)

    # ---- data -----------------------------------------------------------------------------
    def block(self, b: int) -> np.ndarray:
        o, s = int(self.blk_offset[b]), int(self.blk_size[b])
        return self.data[o:o + s].reshape(tuple(int(x) for x in self.blk_shape[b]))

    def to_dense(self) -> np.ndarray:
        """Dense expansion (small tensors only; test helper like the reference tests' GetElem loops)."""
        starts = [np.concatenate([[0], np.cumsum(ix.degs())[:-1]]).astype(int) for ix in self.indexes]
        out = np.zeros([ix.dim for ix in self.indexes], dtype=self.dtype)
        for b in range(self.nblk):
            sl = tuple(slice(starts[i][c], starts[i][c] + int(self.blk_shape[b, i])) for i, c in enumerate(self.blk_coors[b]))
            out[sl] = self.block(b)
        return out

    def norm2(self) -> float:
        return float(np.linalg.norm(self.data))

    def same_structure(self, other: "BlockSparseTensor") -> bool:
        return (self.indexes == other.indexes and np.array_equal(self.blk_coors, other.blk_coors)
                and np.array_equal(self.blk_offset, other.blk_offset))
