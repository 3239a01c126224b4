"""Reader / writer of the reference's tensor file format, so that the contraction path
can be fed real `.qlten` tensors written by TensorToolkit programs, and its results read back by them.

Layout (all header fields are decimal text, one per line; the raw data is binary):
    QLTensor::StreamWrite            qltensor/qltensor_impl.h:823-833      rank, then every Index, then the data tensor
    Index::StreamWrite               qltensor/index.h:182-189              #sectors, every QNSector, direction, dim, hash
    QNSector::StreamWrite            qltensor/qnsct.h:111-115              QN, degeneracy, hash
    <QN>::StreamWrite                qltensor/special_qn/*.h               value(s), hash
    BlockSparseDataTensor::StreamWrite  blk_spar_data_ten/blk_spar_data_ten.h:812-828
                                     #blocks, block coordinates (ascending blk_idx), raw buffer bytes, newline
The readers of the reference take the hashes from the file (they are not recomputed), and Index / QNSector equality
compares hashes, so the writer reproduces every hash function bit for bit (framework/vec_hash.h:12-36 and the
CalcHash_ of each quantum-number type)."""
import io
from typing import List, Union

import numpy as np

from .tensor import BlockSparseTensor, Index, QNKind, QNSector

_M64 = (1 << 64) - 1
_P1, _P2, _P5 = 11400714785074694791, 14029467366897019727, 2870177450012600261


def _u64(v: int) -> int:
    return v & _M64


def _rot(x: int) -> int:
    """_HASH_XXROTATE, framework/vec_hash.h:15"""
    return _u64((x << 31) | (x >> 33))


def qn_hash(kind: QNKind, qn) -> int:
    if kind.name in ("U1QN", "fU1QN"):                       # u1qn.h:139-153, fu1qn.h:125-128
        return _rot(_u64(int(qn[0])))
    if kind.name in ("U1U1QN", "fU1U1QN"):                   # u1u1qn.h:151-174 (fU1U1QN derives from it)
        seg = 1 << 30
        h = _u64(_u64(int(qn[0]) + seg) + _u64(_u64(int(qn[1]) + seg) * (2 * seg)))
        return _u64((h << 10) | (h >> 54))
    if kind.name in ("Z2QN", "fZ2QN"):                       # znqn.h:109-118, fz2qn.h:113-122
        h = _u64(_P5 + _u64(int(qn[0])) * _P2)
        h = _u64(_rot(h) * _P1)
        return _u64(h + (1 ^ _P5))
    raise ValueError(f"no stream format for quantum number type {kind.name}")


def sector_hash(kind: QNKind, s: QNSector) -> int:
    """QNSector::CalcHash_, qnsct.h:138"""
    return qn_hash(kind, s.qn) ^ int(s.dgnc)


def index_hash(ix: Index) -> int:
    """Index::CalcHash_, index.h:225-228: VecHasher over the sectors ^ std::hash<int>(dir) (identity, sign-extended)."""
    h = _P5
    for s in ix.sectors:
        h = _u64(h + sector_hash(ix.kind, s) * _P2)
        h = _u64(_rot(h) * _P1)
    h = _u64(h + (len(ix.sectors) ^ _P5))
    return h ^ _u64(int(ix.dir))


def dumps(t: BlockSparseTensor) -> bytes:
    """The bytes `os << qltensor` writes for this tensor."""
    out = io.BytesIO()
    w = lambda *vals: out.write("".join(f"{int(v)}\n" for v in vals).encode())
    w(t.rank)
    for ix in t.indexes:
        w(ix.nsct)
        for s in ix.sectors:
            w(*s.qn, qn_hash(ix.kind, s.qn), s.dgnc, sector_hash(ix.kind, s))
        w(ix.dir, ix.dim, index_hash(ix))
    # block count = blk_idx_data_blk_map_.size() (blk_spar_data_ten.h:790-796); a scalar never inserts a block, so a
    # rank-0 tensor writes 0 whether or not it holds a value -- only the payload differs
    w(t.nblk if t.rank else 0)
    if t.rank:
        for b in range(t.nblk):
            w(*[int(c) for c in t.blk_coors[b]])
    data = np.ascontiguousarray(t.data)
    if t.rank == 0 and data.size == 0:
        data = np.zeros(1, t.dtype)                           # "empty scalar" case, blk_spar_data_ten.h:820-824
    out.write(data.tobytes())
    out.write(b"\n")
    return out.getvalue()


def save(t: BlockSparseTensor, path: str):
    with open(path, "wb") as f:
        f.write(dumps(t))


class _Tokens:
    """`is >> value` on a byte buffer: skips white space, reads one decimal token."""

    def __init__(self, buf: bytes):
        self.buf, self.pos = buf, 0

    def next_int(self) -> int:
        b, n = self.buf, len(self.buf)
        while self.pos < n and b[self.pos] in b" \t\r\n":
            self.pos += 1
        start = self.pos
        while self.pos < n and b[self.pos] not in b" \t\r\n":
            self.pos += 1
        if start == self.pos:
            raise ValueError("unexpected end of tensor file")
        return int(b[start:self.pos])


def loads(buf: bytes, kind: QNKind, dtype) -> BlockSparseTensor:
    """Inverse of dumps (QLTensor::StreamRead, qltensor_impl.h:799-808).  The file does not name its quantum-number or
    element type -- like the reference's reader, the caller states them."""
    tk = _Tokens(buf)
    rank = tk.next_int()
    indexes: List[Index] = []
    for _ in range(rank):
        nsct = tk.next_int()
        scts = []
        for _ in range(nsct):
            qn = tuple(tk.next_int() for _ in range(kind.nvals))
            tk.next_int()                                      # QN hash
            dg = tk.next_int()
            tk.next_int()                                      # sector hash
            scts.append(QNSector(qn, dg))
        direction = tk.next_int()
        tk.next_int(); tk.next_int()                           # dim, index hash
        indexes.append(Index(kind, scts, direction))
    t = BlockSparseTensor(indexes, dtype)
    nblk = tk.next_int()
    if rank:
        coors = np.array([[tk.next_int() for _ in range(rank)] for _ in range(nblk)], dtype=np.uint32).reshape(nblk, rank)
        if nblk:
            t.set_blocks(coors)
        n = int(t.data.size)
    else:
        n = 1                                                  # scalar: raw_data_size_ = 1 (blk_spar_data_ten.h:805)
    start = tk.pos + 1                                         # RawDataRead_ skips the line break (raw_data_operations.h:585)
    raw = np.frombuffer(buf, dtype=np.dtype(dtype), count=n, offset=start).copy()
    if rank:
        t.data[...] = raw
    else:
        t.data = raw
    return t


def load(path: str, kind: QNKind, dtype) -> BlockSparseTensor:
    with open(path, "rb") as f:
        return loads(f.read(), kind, dtype)
