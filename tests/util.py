"""Shared helpers for the parity tests: seeded random contraction cases in the style of the
reference's fixtures (tests/test_tensor_manipulation/test_ten_ctrct.cc:49-103: U(1) indexes with
3 sectors q=-1,0,+1 of degeneracy 3 or 10), widened to all QN kinds the bridge instantiates."""
import numpy as np

import tensortoolkit_b200 as tk
from tensortoolkit_b200.tensor import IN, OUT, Index, QNSector

KINDS = {"U1": tk.U1, "fU1": tk.fU1, "U1U1": tk.U1U1, "fU1U1": tk.fU1U1, "Z2": tk.Z2, "fZ2": tk.fZ2}


def rel_fro(x, y):
    x = np.asarray(x); y = np.asarray(y)
    d = np.linalg.norm(x - y)
    n = np.linalg.norm(y)
    return d / n if n > 0 else d


def base_index(kind, rng, big=False):
    """A small random index of the given symmetry (OUT direction)."""
    hi = 12 if big else 5
    if kind.name in ("Z2QN", "fZ2QN"):
        qns = [(0,), (1,)]
    elif kind.nvals == 1:
        qns = [(q,) for q in range(-1, 2)] if rng.random() < 0.7 else [(q,) for q in range(-2, 3)]
    else:
        qns = [(0, 0), (1, 1), (1, -1), (2, 0), (-1, 1)]
        qns = [qns[i] for i in sorted(rng.choice(len(qns), size=int(rng.integers(2, len(qns) + 1)), replace=False))]
    return Index(kind, [QNSector(q, int(rng.integers(1, hi))) for q in qns], OUT)


def random_case(kind_name, rng, rank_a=None, rank_b=None, nctrct=None, big=False):
    """Returns (idx_a, idx_b, axes, div_a, div_b) with A[a_i] == InverseIndex(B[b_i])."""
    kind = KINDS[kind_name]
    pool = [base_index(kind, rng, big) for _ in range(3)]
    rank_a = rank_a or int(rng.integers(1, 5))
    rank_b = rank_b or int(rng.integers(1, 5))
    nctrct = min(rank_a, rank_b, int(rng.integers(1, 4))) if nctrct is None else nctrct

    def pick():
        ix = pool[int(rng.integers(len(pool)))]
        return ix if rng.random() < 0.5 else ix.inverse()

    idx_a = [pick() for _ in range(rank_a)]
    idx_b = [pick() for _ in range(rank_b)]
    a_axes = [int(x) for x in rng.choice(rank_a, size=nctrct, replace=False)]
    b_axes = [int(x) for x in rng.choice(rank_b, size=nctrct, replace=False)]
    for x, y in zip(a_axes, b_axes):
        idx_b[y] = idx_a[x].inverse()
    zero = tuple([0] * kind.nvals)
    if kind.name in ("Z2QN", "fZ2QN"):
        divs = [zero, (1,)]
    elif kind.nvals == 1:
        divs = [zero, (1,), (-1,)]
    else:
        divs = [zero, (1, 1), (0, 0)]
    div_a = divs[int(rng.integers(len(divs)))]
    div_b = divs[int(rng.integers(len(divs)))]
    return idx_a, idx_b, (a_axes, b_axes), div_a, div_b


def case_list(n_per_kind=6, seed=20261017, big=False):
    rng = np.random.default_rng(seed)
    cases = []
    for kind_name in KINDS:
        for i in range(n_per_kind):
            dtype = np.float64 if i % 2 == 0 else np.complex128
            cases.append((kind_name, dtype, random_case(kind_name, rng, big=big)))
    return cases


def fixed_cases():
    """The shapes the reference's own contraction tests exercise (test_ten_ctrct.cc:188-612)."""
    out = []
    for kind_name in ("U1", "fU1"):
        kind = KINDS[kind_name]
        for dg in (3, 10):
            i_in = Index(kind, [QNSector((-1,), dg), QNSector((0,), dg), QNSector((1,), dg)], IN)
            i_out = i_in.inverse()
            z = (0,)
            out += [
                (kind_name, "1d", [i_in], [i_out], ([0], [0]), z, z),
                (kind_name, "2d_mm", [i_in, i_out], [i_in, i_out], ([1], [0]), z, z),
                (kind_name, "2d_trace", [i_in, i_out], [i_in, i_out], ([0, 1], [1, 0]), z, z),
                (kind_name, "3d_1axis", [i_in, i_out, i_out], [i_in, i_out, i_out], ([2], [0]), z, z),
                (kind_name, "3d_2axes", [i_in, i_out, i_out], [i_in, i_in, i_out], ([1, 2], [0, 1]), z, z),
                (kind_name, "3d_2axes_trans", [i_in, i_out, i_out], [i_in, i_in, i_out], ([2, 1], [1, 0]), z, z),
                (kind_name, "3d_3axes", [i_in, i_out, i_out], [i_out, i_in, i_in], ([0, 1, 2], [0, 1, 2]), z, z),
                (kind_name, "3d_first_axis", [i_in, i_out, i_out], [i_in, i_in, i_out], ([0], [2]), z, z),
            ]
    return out


def make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, seed):
    ref.set_seed(seed)
    a = ref.RefTensor.new(idx_a, dtype).random(div_a)
    b = ref.RefTensor.new(idx_b, dtype).random(div_b)
    return a, b


def assert_same_as_ref(c_bst, c_ref, tol):
    """Parity protocol of SURVEY.md section 8d: identical indexes and block topology, relative Frobenius <= tol."""
    assert [ix for ix in c_bst.indexes] == list(c_ref.indexes)
    ridx, rcoors, rshape, roff = c_ref.blocks()
    assert c_bst.nblk == len(ridx)
    assert np.array_equal(c_bst.blk_idx, ridx)
    assert np.array_equal(c_bst.blk_coors, rcoors)
    assert np.array_equal(c_bst.blk_shape, rshape)
    assert np.array_equal(c_bst.blk_offset, roff)
    raw = c_ref.raw()
    assert c_bst.data.shape == raw.shape
    if raw.size:
        assert rel_fro(c_bst.data, raw) <= tol


# ---- Python restatement of the row-slab partitioner (the product calls qlb200_shard_* in the library) ----
def cut_line_py(pieces, nsct, degs, world, snap=8):
    degs = [int(d) for d in degs]
    total = sum((hi - lo) * w for _, lo, hi, w in pieces)

    def locate(target):
        acc = 0.0
        for s, lo, hi, w in pieces:
            c = (hi - lo) * w
            if acc + c > target and c > 0:
                r = lo + (target - acc) / w
                r = int(round(r / snap)) * snap
                return s, max(0, min(degs[s], r))
            acc += c
        return nsct, 0
    cuts = [(0, 0)] + [locate(total * r / world) for r in range(1, world)] + [(nsct, 0)]
    for i in range(1, len(cuts)):
        if cuts[i] < cuts[i - 1]:
            cuts[i] = cuts[i - 1]
    out = []
    for r in range(world):
        (s0, r0), (s1, r1) = cuts[r], cuts[r + 1]
        ranges = []
        for s, d in enumerate(degs):
            lo, hi = 0, d
            if s < s0 or s > s1:
                lo = hi = 0
            else:
                if s == s0:
                    lo = r0
                if s == s1:
                    hi = r1
            ranges.append((lo, max(lo, hi)))
        out.append(ranges)
    return out


def reweigh_pieces_py(pieces, ranges_per_rank, times_ms, damp=1.0):
    import numpy as np
    world = len(ranges_per_rank)
    model = []
    for r in range(world):
        w = 0.0
        for s, lo, hi, wt in pieces:
            a, b = max(lo, ranges_per_rank[r][s][0]), min(hi, ranges_per_rank[r][s][1])
            if b > a:
                w += (b - a) * wt
        model.append(w)
    ts = [t for t, m in zip(times_ms, model) if m > 0]
    ms = [m for m in model if m > 0]
    mean_t = (float(np.mean(ts)) if ts else 1.0) or 1.0
    mean_m = (float(np.mean(ms)) if ms else 1.0) or 1.0
    out = []
    for s, lo, hi, wt in pieces:
        for r in range(world):
            a, b = max(lo, ranges_per_rank[r][s][0]), min(hi, ranges_per_rank[r][s][1])
            if b > a:
                f = (times_ms[r] / mean_t) / (model[r] / mean_m) if model[r] > 0 else 1.0
                out.append((s, a, b, wt * f ** damp))
    out.sort(key=lambda p: (p[0], p[1]))
    return out


def restrict_tensor_py(t, axis, ranges):
    """numpy restatement of sharding.restrict_tensor (block-by-block slicing)."""
    import tensortoolkit_b200 as tk
    from tensortoolkit_b200.tensor import Index, QNSector
    ix = t.indexes[axis]
    keep = [s for s, (lo, hi) in enumerate(ranges) if hi > lo]
    new_pos = {s: i for i, s in enumerate(keep)}
    new_ix = Index(ix.kind, [QNSector(ix.sectors[s].qn, ranges[s][1] - ranges[s][0]) for s in keep], ix.dir)
    idxs = list(t.indexes)
    idxs[axis] = new_ix
    out = tk.BlockSparseTensor(idxs, t.dtype)
    sel = [b for b in range(t.nblk) if int(t.blk_coors[b, axis]) in new_pos]
    if not sel or not keep:
        return out
    coors = t.blk_coors[sel].copy()
    coors[:, axis] = [new_pos[int(c)] for c in coors[:, axis]]
    out.set_blocks(coors)
    for nb, b in enumerate(sel):
        lo, hi = ranges[int(t.blk_coors[b, axis])]
        sl = [slice(None)] * t.rank
        sl[axis] = slice(lo, hi)
        out.block(nb)[...] = t.block(b)[tuple(sl)]
    return out
