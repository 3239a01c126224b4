"""Multi-GPU partition of a contraction chain by OUTPUT quantum-number sector and row slab.

The reference distributes the DMRG mat-vec over MPI ranks by restricting one FREE index of the
first operand to a single QN sector per work unit (dmrg::Contract1Sector,
tensor_manipulation/dmrg/contract_1sector.h:181-228; recipe in
tests/test_tensor_manipulation/test_ten_ctrct_1sct.cc:265-279) and summing the partial results.
Because that index stays free through every step of the chain, each unit's intermediates never
leave its rank.  Here the same idea is taken one step further for 8 GPUs behind one NVSwitch:

  * work unit = a RANGE OF ROWS (degeneracy sub-range) of one sector of the split index, so the
    dominant sector (17 % of the flops for the Gaussian bond of the benchmark) can be cut;
  * all rows of the split index form one line, weighted by the flops they cause in every step;
    rank r owns the r-th equal-cost segment of that line (all ranks compute the same cuts);
  * each rank's result blocks are row slabs of the full result's blocks, i.e. contiguous ranges of
    the full raw buffer -- no reduction is needed, only an all-gather of disjoint ranges.
"""
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np

import ctypes as C

from . import _lib
from ._lib import check, lib
from .contract import Match
from .tensor import BlockSparseTensor, Index, QNSector


@dataclass
class Slab:
    full_offset: int     # element offset in the full (unsharded) result raw buffer
    local_offset: int    # element offset in the owning rank's packed result raw buffer
    length: int


@dataclass
class ShardInfo:
    world: int
    rank: int
    sector_ranges: List[List[Tuple[int, int]]]   # [rank][sector] -> (lo, hi) rows of the split index
    slabs: List[List[Slab]]                      # [rank] -> slabs of that rank's result
    local_elems: List[int]                       # [rank] packed result size
    full_elems: int
    cost_share: List[float]                      # [rank] fraction of the chain's flops


def _track_axis(steps, tensors_rank: Dict[str, int], name: str, axis: int):
    """Position of the split index in the lhs operand of every step (it must stay a free lhs axis)."""
    pos = {name: axis}
    out = []
    ranks = dict(tensors_rank)
    for lhs, rhs, axes, res in steps:
        if lhs not in pos:
            raise ValueError("the split index must travel through the lhs operands of the chain")
        p = pos[lhs]
        if p in axes[0]:
            raise ValueError("the split index is contracted inside the chain")
        saved = [i for i in range(ranks[lhs]) if i not in axes[0]]
        out.append(p)
        pos[res] = saved.index(p)
        ranks[res] = len(saved) + (ranks[rhs] - len(axes[1]))
    return out, pos[steps[-1][3]]


def sector_costs(tensors: Dict[str, BlockSparseTensor], steps, name: str, axis: int, dtype) -> Tuple[np.ndarray, Dict[str, BlockSparseTensor], int]:
    """Flops of the whole chain attributed to each sector of the split index."""
    shells = dict(tensors)
    ranks = {k: v.rank for k, v in tensors.items()}
    track, out_axis = _track_axis(steps, ranks, name, axis)
    nsct = tensors[name].indexes[axis].nsct
    cost = np.zeros(nsct)
    code = _lib.C64 if np.dtype(dtype) == np.complex128 else _lib.F64
    for (lhs, rhs, axes, res), p in zip(steps, track):
        m = Match(shells[lhs], shells[rhs], axes)
        check(lib.qlb200_shard_sector_flops(m.h, p, code, cost.ctypes.data_as(C.POINTER(C.c_double))), "qlb200_shard_sector_flops")
        shells[res] = m.result_shell(dtype)
        m.close()
    return cost, shells, out_axis


def line_pieces(cost: np.ndarray, degs: Sequence[int]) -> List[Tuple[int, int, int, float]]:
    """The line of rows of the split index as pieces (sector, lo, hi, weight per row); initially one piece per sector,
    weighted by the flops a row of that sector causes in the whole chain."""
    return [(s, 0, int(d), (float(c) / int(d)) if d else 0.0) for s, (c, d) in enumerate(zip(cost, degs))]


def _pieces_array(pieces):
    arr = (_lib.Piece * max(len(pieces), 1))()
    for i, (s, lo, hi, w) in enumerate(pieces):
        arr[i].sector, arr[i].lo, arr[i].hi, arr[i].weight = int(s), int(lo), int(hi), float(w)
    return arr


def cut_line(pieces, nsct: int, degs: Sequence[int], world: int, snap: int = 8) -> List[List[Tuple[int, int]]]:
    """Cut the line of rows (sector-major) into `world` contiguous segments of equal weight.  Cuts are snapped to
    multiples of `snap` rows inside a sector.  Returns [rank][sector] -> (lo, hi).  (qlb200_shard_cut_line)"""
    d = np.ascontiguousarray(degs, dtype=np.uint32)
    out = np.zeros((world, nsct, 2), np.uint32)
    check(lib.qlb200_shard_cut_line(_pieces_array(pieces), len(pieces), d.ctypes.data_as(C.POINTER(C.c_uint32)), nsct, world, snap,
                                    out.ctypes.data_as(C.POINTER(C.c_uint32))), "qlb200_shard_cut_line")
    return [[(int(lo), int(hi)) for lo, hi in out[r]] for r in range(world)]


def row_line_cuts(cost: np.ndarray, degs: Sequence[int], world: int, snap: int = 8) -> List[List[Tuple[int, int]]]:
    """Cut the line of rows (sector-major) into `world` segments of equal cost."""
    return cut_line(line_pieces(cost, degs), len(degs), degs, world, snap)


def reweigh_pieces(pieces, ranges_per_rank, times_ms: Sequence[float], damp: float = 1.0):
    """Feedback step of the partitioner.  `times_ms[r]` = measured time of rank r's share under the cuts `ranges_per_rank`
    (made from `pieces`).  Time is not proportional to flops -- tile quantisation, small sectors, the memory-bound steps --
    so every rank's rows are re-weighted by (measured time / modelled weight) of that rank; cutting the re-weighted line
    into equal parts moves rows from slow ranks to fast ones.  `damp` < 1 takes only part of the step (factor ** damp): moving
    rows changes tile counts, so the full step tends to overshoot and swap the roles of the ranks.  Returns the new pieces
    (split at the old cuts).  (qlb200_shard_reweigh)"""
    world = len(ranges_per_rank)
    nsct = len(ranges_per_rank[0]) if world else 0
    rg = np.ascontiguousarray(ranges_per_rank, dtype=np.uint32).reshape(world, nsct, 2)
    t = np.ascontiguousarray(times_ms, dtype=np.float64)
    arr = _pieces_array(pieces)
    args = (arr, len(pieces), rg.ctypes.data_as(C.POINTER(C.c_uint32)), nsct, world, t.ctypes.data_as(C.POINTER(C.c_double)), float(damp))
    n = int(lib.qlb200_shard_reweigh(*args, 0, None))
    out = (_lib.Piece * max(n, 1))()
    lib.qlb200_shard_reweigh(*args, n, out)
    return [(int(out[i].sector), int(out[i].lo), int(out[i].hi), float(out[i].weight)) for i in range(n)]


def tune_partition(chain, build, measure, rounds: int = 4, damp: float = 0.5):
    """Partition feedback loop (set-up time autotuning, like the planner's split-K simulation): `chain` is the sharded chain
    built on the flop-balanced cuts (anything with `.info.pieces`, `.info.sector_ranges` and `.close()`), `build(pieces)` builds one
    on other weights, `measure(chain)` returns the local time of EVERY rank (the same list on all ranks, e.g. all-gathered), so
    all ranks take the same decisions without talking.  Every candidate cut is measured; the one whose slowest rank is fastest
    is kept (a step can overshoot: moving rows changes tile counts).  Returns (chain, index of the kept cut, measured times of
    every cut)."""
    log, best = [], None
    for it in range(rounds + 1):
        times = [float(t) for t in measure(chain)]
        log.append(times)
        if best is None or max(times) < best[0]:
            best = (max(times), chain.info.pieces, it)
        if it == rounds:
            break
        pieces = reweigh_pieces(chain.info.pieces, chain.info.sector_ranges, times, damp=damp)
        chain.close()
        chain = build(pieces)
    if best[2] != rounds:
        chain.close()
        chain = build(best[1])
    return chain, best[2], log


def slab_layout(t: BlockSparseTensor, axis: int, ranges: Sequence[Tuple[int, int]]):
    """qlb200_shard_restrict: (kept old sector numbers, their new degeneracies, kept old block ordinals, the blocks' coordinates
    in the slab, (src, dst, len) element ranges that build the slab's raw buffer from the full one)."""
    sv = t.shell()
    rg = np.ascontiguousarray(ranges, dtype=np.uint32).reshape(-1)
    info = _lib.SlabInfo()
    u32, u64 = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    rgp = rg.ctypes.data_as(u32)
    check(lib.qlb200_shard_restrict(sv.ptr(), axis, rgp, C.byref(info), None, None, None, None, None, None, None), "qlb200_shard_restrict")
    ks, nd = np.zeros(max(info.nsct_kept, 1), np.uint32), np.zeros(max(info.nsct_kept, 1), np.uint32)
    kb, nc = np.zeros(max(info.nblk_kept, 1), np.uint32), np.zeros((max(info.nblk_kept, 1), max(t.rank, 1)), np.uint32)
    cs, cd, cl = (np.zeros(max(info.ncopy, 1), np.uint64) for _ in range(3))
    check(lib.qlb200_shard_restrict(sv.ptr(), axis, rgp, C.byref(info), ks.ctypes.data_as(u32), nd.ctypes.data_as(u32), kb.ctypes.data_as(u32),
                                    nc.ctypes.data_as(u32), cs.ctypes.data_as(u64), cd.ctypes.data_as(u64), cl.ctypes.data_as(u64)),
          "qlb200_shard_restrict")
    n = int(info.ncopy)
    return ks[:info.nsct_kept], nd[:info.nsct_kept], kb[:info.nblk_kept], nc[:info.nblk_kept], (cs[:n], cd[:n], cl[:n]), int(info.elems)


def restrict_tensor(t: BlockSparseTensor, axis: int, ranges: Sequence[Tuple[int, int]]) -> BlockSparseTensor:
    """Sub-tensor keeping rows [lo, hi) of every sector of index `axis` (empty sectors dropped): structure and copy list from
    the library (qlb200_shard_restrict), applied here on the host buffer."""
    ix = t.indexes[axis]
    kept, new_deg, blocks, coors, (src, dst, ln), elems = slab_layout(t, axis, ranges)
    new_ix = Index(ix.kind, [QNSector(ix.sectors[int(s)].qn, int(d)) for s, d in zip(kept, new_deg)], ix.dir)
    idxs = list(t.indexes)
    idxs[axis] = new_ix
    out = BlockSparseTensor(idxs, t.dtype)
    if not len(blocks) or not len(kept):
        return out
    out.set_blocks(coors)          # relabelling is monotone, so block order is preserved
    assert out.data.size == elems
    for s0, d0, n in zip(src.tolist(), dst.tolist(), ln.tolist()):
        out.data[d0:d0 + n] = t.data[s0:s0 + n]
    return out


def shard_chain(tensors: Dict[str, BlockSparseTensor], steps, name: str, axis: int, world: int, rank: int,
                dtype=None, pieces=None, snap: int = 8) -> Tuple[Dict[str, BlockSparseTensor], ShardInfo]:
    """Returns this rank's operand set (only `name` differs) and the slab map of every rank.  `pieces`: a re-weighted row
    line from reweigh_pieces (measured feedback); default = rows weighted by their flops."""
    dtype = dtype or tensors[name].dtype
    cost, shells, out_axis = sector_costs(tensors, steps, name, axis, dtype)
    if out_axis != 0:
        raise ValueError("the split index must end up as the first index of the result (row slabs must be contiguous)")
    degs = tensors[name].indexes[axis].degs()
    if pieces is None:
        pieces = line_pieces(cost, degs)
    ranges = cut_line(pieces, len(degs), degs, world, snap)
    full_out = shells[steps[-1][3]]
    per_row = [c / d if d else 0.0 for c, d in zip(cost, degs)]
    total = float(cost.sum()) or 1.0
    slabs, local_elems, share = [], [], []
    for r in range(world):
        lst, loc = [], 0
        for b in range(full_out.nblk):           # ascending blk_idx == the rank's own packed block order
            s = int(full_out.blk_coors[b, 0])
            lo, hi = ranges[r][s]
            if hi <= lo:
                continue
            rest = int(full_out.blk_size[b]) // int(full_out.blk_shape[b, 0])
            lst.append(Slab(int(full_out.blk_offset[b]) + lo * rest, loc, (hi - lo) * rest))
            loc += (hi - lo) * rest
        slabs.append(lst)
        local_elems.append(loc)
        share.append(sum(per_row[s] * (hi - lo) for s, (lo, hi) in enumerate(ranges[r])) / total)
    mine = dict(tensors)
    mine[name] = restrict_tensor(tensors[name], axis, ranges[rank])
    info = ShardInfo(world, rank, ranges, slabs, local_elems, int(full_out.data.size), share)
    info.pieces = pieces
    return mine, info


def shard_heff_tensors(tensors, world, rank):
    """Two-site H_eff apply (workloads.HEFF_STEPS): split lenv's free ket-side bond (axis 2)."""
    from .workloads import HEFF_STEPS
    return shard_chain(tensors, HEFF_STEPS, "lenv", 2, world, rank)


def unpack_slabs(info: ShardInfo, gathered: np.ndarray, stride: int, full: np.ndarray):
    """Host model of the unpack step: gathered[r*stride + local] -> full[full_offset] (tests)."""
    for r in range(info.world):
        for s in info.slabs[r]:
            full[s.full_offset:s.full_offset + s.length] = gathered[r * stride + s.local_offset:r * stride + s.local_offset + s.length]
