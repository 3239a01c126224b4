"""qlten::ContractContiguousAxes: the contiguous-axes contraction of the reference
(tensor_manipulation/contract_contiguous_axes.h:849-873) behind the same matcher / plan / grouped GEMM.

Host part (no GPU): the numpy restatement against the reference run here for all four CtrctSide pairs, our matcher's
block structure / task signs against both, and the plan reading every block in place.  GPU part: values against the
reference, through the Python API and through the C++ drop-in adapter, and the reference's own property test
(test_ten_ctrct.cc:705-728: the contiguous variant equals Contract followed by Transpose)."""
import numpy as np
import pytest

import tensortoolkit_b200 as tk
from tensortoolkit_b200 import _lib
from oracle import contract_np as onp
from tests import util

TOL = 1e-12
SIDES = [("tail", "head"), ("head", "head"), ("tail", "tail"), ("head", "tail")]


def contiguous_case(kind_name, rng, big=False):
    """Random operands whose contracted axes are cyclically contiguous: A[(a_start+i) % ra] == Inverse(B[(b_start+i) % rb])."""
    kind = util.KINDS[kind_name]
    pool = [util.base_index(kind, rng, big) for _ in range(3)]
    ra, rb = int(rng.integers(1, 6)), int(rng.integers(1, 6))
    size = int(rng.integers(0 if min(ra, rb) > 1 else 1, min(ra, rb) + 1))
    size = max(size, 1) if ra + rb - 2 * size >= 0 else size
    a_start, b_start = int(rng.integers(ra)), int(rng.integers(rb))

    def pick():
        ix = pool[int(rng.integers(len(pool)))]
        return ix if rng.random() < 0.5 else ix.inverse()

    idx_a = [pick() for _ in range(ra)]
    idx_b = [pick() for _ in range(rb)]
    for i in range(size):
        idx_b[(b_start + i) % rb] = idx_a[(a_start + i) % ra].inverse()
    zero = tuple([0] * kind.nvals)
    if kind.name in ("Z2QN", "fZ2QN"):
        divs = [zero, (1,)]
    elif kind.nvals == 1:
        divs = [zero, (1,), (-1,)]
    else:
        divs = [zero, (1, 1), (0, 0)]
    return idx_a, idx_b, (a_start, b_start, size), divs[int(rng.integers(len(divs)))], divs[int(rng.integers(len(divs)))]


def case_list(n_per_kind, seed, big=False):
    rng = np.random.default_rng(seed)
    out = []
    for kind_name in util.KINDS:
        for i in range(n_per_kind):
            out.append((kind_name, np.float64 if i % 2 == 0 else np.complex128, contiguous_case(kind_name, rng, big)))
    return out


CASES = case_list(8, 20261018)
BIG = case_list(2, 20261019, big=True)


@pytest.mark.parametrize("case", range(len(CASES)))
def test_oracle_matches_reference_for_every_side_pair(ref, case):
    """The restatement is pinned by the reference itself; all four CtrctSide pairs give the same tensor."""
    kind_name, dtype, (idx_a, idx_b, (a0, b0, n), div_a, div_b) = CASES[case]
    a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, 900 + case)
    want = ref.contract_contiguous(a, b, a0, b0, n)
    got = onp.contract_contiguous_np(a.to_bst(), b.to_bst(), a0, b0, n)
    util.assert_same_as_ref(got, want, 1e-13)
    for sides in SIDES[1:]:
        other = ref.contract_contiguous(a, b, a0, b0, n, sides)
        assert list(other.indexes) == list(want.indexes)
        assert all(np.array_equal(x, y) for x, y in zip(other.blocks(), want.blocks()))
        assert util.rel_fro(other.raw(), want.raw()) <= 1e-13


@pytest.mark.parametrize("case", range(len(CASES)))
def test_matcher_structure_and_signs(ref, case):
    """qlb200_match_create_contiguous: result indexes, block map (keys, coordinates, shapes, offsets) identical to
    the reference's output; task list (pairs, m/k/n, offsets, signs incl. the residue signs) identical to the oracle's."""
    kind_name, dtype, (idx_a, idx_b, (a0, b0, n), div_a, div_b) = CASES[case]
    a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, 900 + case)
    A, B = a.to_bst(), b.to_bst()
    want = ref.contract_contiguous(a, b, a0, b0, n)
    m = tk.Match(A, B, None, contiguous=(a0, b0, n))
    assert list(m.c_indexes) == list(want.indexes)
    ra, rb = len(idx_a), len(idx_b)
    assert m.saved_axes == ([(a0 + n + i) % ra for i in range(ra - n)], [(b0 + n + i) % rb for i in range(rb - n)])
    if m.c_rank > 0:
        idx, coors, shape, off = m.c_blocks()
        ridx, rcoors, rshape, roff = want.blocks()
        assert np.array_equal(idx, ridx) and np.array_equal(coors, rcoors)
        assert np.array_equal(shape, rshape) and np.array_equal(off, roff)
    assert m.c_elems == want.raw().size
    # tasks vs the oracle's restatement of GenerateDataBlk_ + residue signs
    axes = ([(a0 + i) % ra for i in range(n)], [(b0 + i) % rb for i in range(n)])
    sa, sb = m.saved_axes
    ot, _, _ = onp.match_tasks(A, B, axes, saved=(sa, sb))
    a_end = (a0 + n) % ra
    got = m.tasks()
    assert len(got) == len(ot)
    for g, o in zip(got, ot):
        assert (g.a_blk_idx, g.b_blk_idx, g.c_blk_idx, g.a_off, g.b_off, g.c_off, g.m, g.k, g.n) == \
               (o["a_idx"], o["b_idx"], o["c_idx"], o["a_off"], o["b_off"], o["c_off"], o["m"], o["k"], o["n"])
        sign = o["sign"]
        if A.kind.fermionic:
            if a_end > 0:
                sign *= onp.residue_fermion_sign(onp._blk_parities(A, o["a"]), sa, a_end)
            if b0 > 0:
                sign *= onp.residue_fermion_sign(onp._blk_parities(B, o["b"]), sb, b0)
        assert g.sign == sign and bool(g.first) == (o["beta"] == 0.0)
    m.close()


def test_precondition_errors():
    rng = np.random.default_rng(3)
    idx_a, idx_b, (a0, b0, n), div_a, div_b = contiguous_case("U1", rng)
    A = tk.BlockSparseTensor(idx_a, np.float64).random(div_a, rng)
    B = tk.BlockSparseTensor(idx_b, np.float64).random(div_b, rng)
    with pytest.raises(ValueError):
        tk.Match(A, B, None, contiguous=(len(idx_a), b0, n))
    with pytest.raises(ValueError):
        tk.Match(A, B, None, contiguous=(a0, b0, min(len(idx_a), len(idx_b)) + 1))


def test_heff_chain_contiguous_plans_read_blocks_in_place():
    """The DMRG pattern the API exists for (tail of A with head of B): no block goes through the permute kernel, and
    neither do the rotated variants -- a cyclic rotation is a 2-D transposition the GEMM producers absorb."""
    from tensortoolkit_b200 import workloads as wl
    rng = np.random.default_rng(4)
    ix = wl.u1_heisenberg_indexes(64)
    psi = tk.BlockSparseTensor([ix["vb_in"], ix["ph_out"], ix["ph_out"], ix["vb_out"]], np.complex128).random((0,), rng)
    renv = tk.BlockSparseTensor([ix["vb_in"], ix["wb_in"], ix["vb_out"]], np.complex128).random((0,), rng)
    lenv = tk.BlockSparseTensor([ix["vb_out"], ix["wb_out"], ix["vb_in"]], np.complex128).random((0,), rng)
    for a, b, spec in ((psi, renv, (3, 0, 1)), (lenv, psi, (0, 0, 1)), (psi, lenv, (0, 0, 1)), (renv, psi, (2, 0, 1))):
        m = tk.Match(a, b, None, contiguous=spec)
        assert m.ntask > 0
        p = tk.ContractionPlan(None, m, np.complex128)
        st = p.stats()
        assert st.permute_elems_a == 0 and st.permute_elems_b == 0
        p.close(); m.close()


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(CASES)))
def test_contract_contiguous_axes_vs_reference(ref, ctx, case):
    kind_name, dtype, (idx_a, idx_b, (a0, b0, n), div_a, div_b) = CASES[case]
    a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, 900 + case)
    got = tk.contract_contiguous_axes(a.to_bst(), b.to_bst(), a0, b0, n, ctx)
    util.assert_same_as_ref(got, ref.contract_contiguous(a, b, a0, b0, n), TOL)


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(BIG)))
def test_bigger_blocks_and_generic_variant_property(ref, ctx, case):
    """Reference property (test_ten_ctrct.cc:705-728): the contiguous variant == Contract + Transpose to the cyclic order."""
    kind_name, dtype, (idx_a, idx_b, (a0, b0, n), div_a, div_b) = BIG[case]
    a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, 950 + case)
    A, B = a.to_bst(), b.to_bst()
    got = tk.contract_contiguous_axes(A, B, a0, b0, n, ctx)
    util.assert_same_as_ref(got, ref.contract_contiguous(a, b, a0, b0, n), TOL)
    ra, rb = len(idx_a), len(idx_b)
    axes = ([(a0 + i) % ra for i in range(n)], [(b0 + i) % rb for i in range(n)])
    generic = tk.contract(A, B, axes, ctx)
    sa = [i for i in range(ra) if i not in axes[0]]
    sb = [i for i in range(rb) if i not in axes[1]]
    cyc = [(a0 + n + i) % ra for i in range(ra - n)] + [ra + (b0 + n + i) % rb for i in range(rb - n)]
    order = [(sa + [ra + x for x in sb]).index(ax) for ax in cyc]
    if generic.rank > 1 and generic.nblk:
        moved = tk.transpose(generic, order, ctx)
        assert np.array_equal(moved.blk_coors, got.blk_coors)
        if not A.kind.fermionic:       # for fermions the two routes differ by the documented residue signs
            assert util.rel_fro(moved.data, got.data) <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(0, len(CASES), 3))
def test_dropin_adapter_contiguous(ref, ctx, case):
    """qlten::b200::ContractContiguousAxes<ASide, BSide>(a, b, a_start, b_start, size, c) on reference QLTensors."""
    kind_name, dtype, (idx_a, idx_b, (a0, b0, n), div_a, div_b) = CASES[case]
    a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, 900 + case)
    want = ref.contract_contiguous(a, b, a0, b0, n)
    for sides in (SIDES[0], SIDES[case % 4]):
        got = ref.b200_contract_contiguous(a, b, a0, b0, n, sides, ctx.h)
        assert list(got.indexes) == list(want.indexes)
        assert all(np.array_equal(x, y) for x, y in zip(got.blocks(), want.blocks()))
        if want.raw().size:
            assert util.rel_fro(got.raw(), want.raw()) <= TOL


# ------------------------------------------------------------------------------------------------ golden vectors
import glob
import os

from tests.golden import io as gio

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "contig_*.npz")))


def test_golden_present():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize("path", GOLDEN)
def test_oracle_matches_golden(path):
    """Fixtures dumped from the reference's ContractContiguousAxes (tests/golden/make_golden.py); no reference needed."""
    g = gio.load_case(path)
    a0, b0, n = g["meta"]["contiguous"]
    c = onp.contract_contiguous_np(g["A"], g["B"], a0, b0, n)
    assert c.indexes == g["C"].indexes
    assert np.array_equal(c.blk_coors, g["C"].blk_coors) and np.array_equal(c.blk_offset, g["C"].blk_offset)
    assert util.rel_fro(c.data, g["C"].data) <= 1e-13
    m = tk.Match(g["A"], g["B"], None, contiguous=(a0, b0, n))
    assert m.c_indexes == g["C"].indexes and m.c_elems == g["C"].data.size
    _, coors, _, off = m.c_blocks()
    assert np.array_equal(coors, g["C"].blk_coors) and np.array_equal(off, g["C"].blk_offset)
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN)
def test_gpu_matches_golden(ctx, path):
    g = gio.load_case(path)
    a0, b0, n = g["meta"]["contiguous"]
    c = tk.contract_contiguous_axes(g["A"], g["B"], a0, b0, n, ctx)
    assert c.same_structure(g["C"]) and c.indexes == g["C"].indexes
    assert util.rel_fro(c.data, g["C"].data) <= TOL
