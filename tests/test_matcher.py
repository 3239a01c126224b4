"""Host sector matcher (qlb200_match_*) against the reference's own task list and block structure.
No GPU needed: the matcher is host code behind the C ABI."""
import numpy as np
import pytest

import tensortoolkit_b200 as tk
from oracle import contract_np as onp
from tests import util

CASES = util.case_list(n_per_kind=6)
FIXED = util.fixed_cases()


def tasks_as_arrays(match, sorted_by_c=False):
    t = match.tasks(sorted_by_c)
    u = np.array([[x.a_blk_idx, x.b_blk_idx, x.c_blk_idx, x.a_off, x.b_off, x.c_off, x.m, x.k, x.n] for x in t], dtype=np.uint64).reshape(-1, 9)
    d = np.array([[float(x.sign), 0.0 if x.first else 1.0] for x in t], dtype=np.float64).reshape(-1, 2)
    return u, d


def check_against_ref(ref, idx_a, idx_b, axes, div_a, div_b, dtype, seed):
    a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, seed)
    A, B = a.to_bst(), b.to_bst()
    m = tk.Match(A, B, axes)
    ru, rd = ref.contract_tasks(a, b, axes, sorted_by_c=False)
    u, d = tasks_as_arrays(m)
    assert np.array_equal(u, ru), "task table differs from RawDataCtrctTask list"
    assert np.array_equal(d, rd), "fermion signs / beta flags differ"
    # sorted order: same multiset per C block, beta=0 task first (reference sort is not stable)
    su, sd = tasks_as_arrays(m, sorted_by_c=True)
    rsu, rsd = ref.contract_tasks(a, b, axes, sorted_by_c=True)
    assert np.array_equal(su[:, 2], rsu[:, 2])
    for c in np.unique(su[:, 2]):
        sel, rsel = su[:, 2] == c, rsu[:, 2] == c
        assert sd[sel][0, 1] == 0.0 and rsd[rsel][0, 1] == 0.0
        assert sorted(map(tuple, su[sel])) == sorted(map(tuple, rsu[rsel]))
    # output block structure
    c_ref = ref.contract(a, b, axes)
    if m.c_rank > 0:
        idx, coors, shape, off = m.c_blocks()
        ridx, rcoors, rshape, roff = c_ref.blocks()
        assert np.array_equal(idx, ridx) and np.array_equal(coors, rcoors)
        assert np.array_equal(shape, rshape) and np.array_equal(off, roff)
        assert m.c_elems == c_ref.raw().size
    else:
        assert m.is_scalar and m.c_elems == c_ref.raw().size
    # cost model == EstimateContractCost
    rc = ref.contract_cost(a, b, axes)
    c = m.cost(dtype)
    for k in ("flops", "gemm_count", "candidate_block_pair_count", "read_bytes", "write_bytes", "temp_peak_bytes"):
        assert float(getattr(c, k)) == rc[k], k
    if m.c_rank > 0:
        assert c.output_block_count == rc["output_block_count"] and c.output_raw_elem_count == rc["output_raw_elem_count"]
    # the numpy oracle restates the same algorithm
    ot, oc, oe = onp.match_tasks(A, B, axes)
    assert [(t["a_idx"], t["b_idx"], t["c_idx"], t["a_off"], t["b_off"], t["c_off"], t["m"], t["k"], t["n"]) for t in ot] == [tuple(int(v) for v in r) for r in ru]
    assert [(float(t["sign"]), t["beta"]) for t in ot] == [tuple(r) for r in rd]
    m.close()


@pytest.mark.parametrize("case", range(len(CASES)))
def test_random_cases_match_reference(ref, case):
    kind_name, dtype, (idx_a, idx_b, axes, div_a, div_b) = CASES[case]
    check_against_ref(ref, idx_a, idx_b, axes, div_a, div_b, dtype, 1000 + case)


@pytest.mark.parametrize("case", range(len(FIXED)))
def test_reference_fixture_shapes(ref, case):
    kind_name, name, idx_a, idx_b, axes, div_a, div_b = FIXED[case]
    check_against_ref(ref, idx_a, idx_b, axes, div_a, div_b, np.float64, 7)


def test_appendix_d_golden():
    """SURVEY.md Appendix D, dumped from the reference: U1, Contract(A,B,{{2,1},{1,0}})."""
    from tensortoolkit_b200.tensor import IN, Index, QNSector
    i_in = Index(tk.U1, [QNSector((-1,), 2), QNSector((0,), 3), QNSector((1,), 2)], IN)
    i_out = i_in.inverse()
    A = tk.BlockSparseTensor([i_in, i_out, i_out]); A.set_blocks(A.div_blocks((0,)))
    B = tk.BlockSparseTensor([i_in, i_in, i_out]); B.set_blocks(B.div_blocks((0,)))
    assert list(A.blk_idx) == [1, 3, 11, 13, 15, 23, 25] and list(A.blk_offset) == [0, 12, 24, 36, 63, 75, 87]
    assert list(B.blk_idx) == [3, 7, 9, 13, 17, 19, 23] and A.data.size == 99 and B.data.size == 99
    m = tk.Match(A, B, ([2, 1], [1, 0]))
    assert m.perm(0) == ([0, 2, 1], True) and m.perm(1) == ([1, 0, 2], True)
    got = [(t.a_blk_idx, t.a_off, t.b_blk_idx, t.b_off, t.c_blk_idx, t.c_off, t.m, t.k, t.n, t.sign, t.first) for t in m.tasks(True)]
    want = [(1, 0, 3, 0, 0, 0, 2, 6, 2, 1, 1), (3, 12, 9, 24, 0, 0, 2, 6, 2, 1, 0),
            (11, 24, 7, 12, 4, 4, 3, 4, 3, 1, 1), (13, 36, 13, 36, 4, 4, 3, 9, 3, 1, 0), (15, 63, 19, 75, 4, 4, 3, 4, 3, 1, 0),
            (23, 75, 17, 63, 8, 13, 2, 6, 2, 1, 1), (25, 87, 23, 87, 8, 13, 2, 6, 2, 1, 0)]
    assert got == want
    idx, coors, shape, off = m.c_blocks()
    assert list(idx) == [0, 4, 8] and list(off) == [0, 4, 13] and m.c_elems == 17
    assert shape.tolist() == [[2, 2], [3, 3], [2, 2]]


def test_cost_known_answer():
    """tests/test_tensor_manipulation/test_tensor_op_cost.cc:19-43: 2x3 . 3x4 -> 48 flops, 144 B read, 64 B written."""
    from tensortoolkit_b200.tensor import IN, OUT, Index, QNSector
    i2 = Index(tk.U1, [QNSector((0,), 2)], OUT)
    i3 = Index(tk.U1, [QNSector((0,), 3)], OUT)
    i4 = Index(tk.U1, [QNSector((0,), 4)], OUT)
    A = tk.BlockSparseTensor([i2, i3]); A.set_blocks(A.div_blocks((0,)))
    B = tk.BlockSparseTensor([i3.inverse(), i4]); B.set_blocks(B.div_blocks((0,)))
    m = tk.Match(A, B, ([1], [0]))
    c = m.cost(np.float64)
    assert c.flops == 48.0 and c.read_bytes == 18 * 8 and c.write_bytes == 8 * 8 and c.gemm_count == 1


def test_one_sector_matches_reference(ref):
    rng = np.random.default_rng(5)
    for kind_name in ("U1", "fU1U1"):
        idx_a, idx_b, axes, div_a, div_b = util.random_case(kind_name, rng, rank_a=3, rank_b=3, nctrct=1)
        a, b = util.make_ref_pair(ref, idx_a, idx_b, np.float64, div_a, div_b, 11)
        A, B = a.to_bst(), b.to_bst()
        free = [i for i in range(3) if i not in axes[0]][0]
        for s in range(idx_a[free].nsct):
            m = tk.Match(A, B, axes, one_sector=(free, s))
            c_ref = ref.contract_1sector(a, free, s, b, axes)
            idx, coors, shape, off = m.c_blocks()
            ridx, rcoors, rshape, roff = c_ref.blocks()
            assert np.array_equal(idx, ridx) and np.array_equal(off, roff) and np.array_equal(shape, rshape)
            m.close()


def test_precondition_errors():
    from tensortoolkit_b200.tensor import IN, Index, QNSector
    i_in = Index(tk.U1, [QNSector((0,), 2), QNSector((1,), 2)], IN)
    A = tk.BlockSparseTensor([i_in, i_in.inverse()]); A.set_blocks(A.div_blocks((0,)))
    with pytest.raises(ValueError):
        tk.Match(A, A, ([0], [0]))          # IN with IN: not inverse indexes
    with pytest.raises(ValueError):
        tk.Match(A, A, ([0, 1], [1]))


def test_empty_operand_gives_no_tasks():
    from tensortoolkit_b200.tensor import IN, Index, QNSector
    i_in = Index(tk.U1, [QNSector((0,), 2), QNSector((1,), 2)], IN)
    A = tk.BlockSparseTensor([i_in, i_in.inverse()])            # no stored blocks
    B = tk.BlockSparseTensor([i_in, i_in.inverse()]); B.set_blocks(B.div_blocks((0,)))
    m = tk.Match(A, B, ([1], [0]))
    assert m.ntask == 0 and m.c_nblk == 0 and m.c_elems == 0


def _task_rows(m):
    return [(x.a_blk_idx, x.b_blk_idx, x.c_blk_idx, x.a_off, x.b_off, x.c_off, x.a_ord, x.b_ord, x.c_ord, x.m, x.k, x.n, x.sign, x.first)
            for x in m.tasks()]


def test_large_sector_space_paths_agree_with_dense_tables(monkeypatch):
    """The matcher keeps dense tables (contracted-key buckets, C block bitmap) when the sector spaces are small and
    falls back to binary search / hashing otherwise; QLB200_MATCH_NO_DENSE forces the fallbacks.  Same tasks, same C
    blocks, on random cases of every symmetry (general and contiguous-axes matching) and on the fermionic H_eff chain."""
    from tensortoolkit_b200 import workloads as wl
    rng = np.random.default_rng(99)
    pairs = []
    for kind_name in util.KINDS:
        for _ in range(6):
            idx_a, idx_b, axes, div_a, div_b = util.random_case(kind_name, rng)
            A = tk.BlockSparseTensor(idx_a, np.float64).random(div_a, rng)
            B = tk.BlockSparseTensor(idx_b, np.float64).random(div_b, rng)
            pairs.append((A, B, axes))
    ti = wl.heff_tensor_indexes(wl.hubbard_indexes(60))
    t = {n: tk.BlockSparseTensor(ix, np.float64).random((0, 0), rng) for n, ix in ti.items()}
    for lhs, rhs, axes, out in wl.HEFF_STEPS:
        pairs.append((t[lhs], t[rhs], axes))
        m = tk.Match(t[lhs], t[rhs], axes); t[out] = m.result_shell(np.float64); m.close()
    for A, B, axes in pairs:
        monkeypatch.delenv("QLB200_MATCH_NO_DENSE", raising=False)
        dense = tk.Match(A, B, axes)
        monkeypatch.setenv("QLB200_MATCH_NO_DENSE", "1")
        sparse = tk.Match(A, B, axes)
        assert _task_rows(dense) == _task_rows(sparse)
        assert dense.c_elems == sparse.c_elems and dense.c_nblk == sparse.c_nblk
        if dense.c_rank:
            assert all(np.array_equal(x, y) for x, y in zip(dense.c_blocks(), sparse.c_blocks()))
        dense.close(); sparse.close()
    monkeypatch.delenv("QLB200_MATCH_NO_DENSE", raising=False)
