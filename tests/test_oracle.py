"""Pins the numpy restatement (oracle/contract_np.py) to the reference: against the reference
library itself when it is built, and against golden vectors dumped from it (tests/golden/)."""
import glob
import os

import numpy as np
import pytest

import tensortoolkit_b200 as tk
from oracle import contract_np as onp
from tests import util
from tests.golden import io as gio

CASES = util.case_list(n_per_kind=4, seed=77)
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith(("contig_", "file_")))     # those belong to test_contiguous / test_qlten_io


@pytest.mark.parametrize("case", range(len(CASES)))
def test_contract_np_matches_reference(ref, case):
    kind_name, dtype, (idx_a, idx_b, axes, div_a, div_b) = CASES[case]
    a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, 300 + case)
    c = onp.contract_np(a.to_bst(), b.to_bst(), axes)
    util.assert_same_as_ref(c, ref.contract(a, b, axes), 1e-13)


@pytest.mark.parametrize("case", range(0, len(CASES), 3))
def test_transpose_np_matches_reference(ref, case):
    kind_name, dtype, (idx_a, _, _, div_a, _) = CASES[case]
    ref.set_seed(case)
    a = ref.RefTensor.new(idx_a, dtype).random(div_a)
    rng = np.random.default_rng(case)
    perm = [int(x) for x in rng.permutation(len(idx_a))]
    got = onp.transpose_np(a.to_bst(), perm)
    want = a.clone().transpose(perm)
    util.assert_same_as_ref(got, want, 0.0)


def test_fermion_sign_known_answers():
    """tests/test_utility/test_fermion_parity_exchange.cc: reorder signs; and the trace convention of
    test_fermion_ten_ctrct.cc:197-210 (odd sector contracted IN->OUT gives -1)."""
    assert onp.fermionic_reorder_sign([1, 1], [1, 0]) == -1
    assert onp.fermionic_reorder_sign([1, 0, 1], [2, 1, 0]) == -1
    assert onp.fermionic_reorder_sign([0, 1, 0], [2, 1, 0]) == 1
    assert onp.fermion_exchange_sign([1, 1], [1, 1], [0, 1], [1, 0], [-1, 1]) in (-1, 1)
    # <bra|ket> adjacent OUT-IN pair: no sign; IN index of A contracted with odd parity: -1
    assert onp.fermion_exchange_sign([1], [1], [0], [0], [1]) == 1
    assert onp.fermion_exchange_sign([1], [1], [0], [0], [-1]) == -1


@pytest.mark.parametrize("path", GOLDEN)
def test_golden_vectors(path):
    g = gio.load_case(path)
    c = onp.contract_np(g["A"], g["B"], g["axes"])
    assert np.array_equal(c.blk_coors, g["C"].blk_coors) and np.array_equal(c.blk_offset, g["C"].blk_offset)
    assert util.rel_fro(c.data, g["C"].data) <= 1e-13
    tasks, _, _ = onp.match_tasks(g["A"], g["B"], g["axes"])
    got = np.array([[t["a_idx"], t["b_idx"], t["c_idx"], t["a_off"], t["b_off"], t["c_off"], t["m"], t["k"], t["n"]] for t in tasks], dtype=np.uint64).reshape(-1, 9)
    assert np.array_equal(got, g["tasks_u"])
    assert np.array_equal(np.array([[t["sign"], t["beta"]] for t in tasks], dtype=np.float64).reshape(-1, 2), g["tasks_d"])


def test_golden_present():
    assert len(GOLDEN) >= 6
