"""The reference's tensor file format: tensortoolkit_b200/qlten_io.py against files written and
read by the reference itself (oracle/_ref), byte for byte.  Host only."""
import numpy as np
import pytest

import tensortoolkit_b200 as tk
from tensortoolkit_b200 import qlten_io
from tests import util

CASES = util.case_list(n_per_kind=4, seed=4242)


@pytest.mark.parametrize("case", range(len(CASES)))
def test_round_trip_through_the_reference(ref, tmp_path, case):
    kind_name, dtype, (idx_a, _, _, div_a, _) = CASES[case]
    ref.set_seed(500 + case)
    a = ref.RefTensor.new(idx_a, dtype).random(div_a)
    A = a.to_bst()
    # (1) reference writes, we read: same indexes, block map, data
    f_ref = tmp_path / "ref.qlten"
    a.write_file(f_ref)
    got = qlten_io.load(str(f_ref), util.KINDS[kind_name], dtype)
    assert got.indexes == A.indexes and got.same_structure(A)
    assert np.array_equal(got.data, A.data)
    # (2) we write: byte-identical to the reference's file (every hash included)
    assert qlten_io.dumps(A) == f_ref.read_bytes()
    # (3) the reference reads our file: indexes compare equal (hash-based), data identical
    f_ours = tmp_path / "ours.qlten"
    qlten_io.save(A, str(f_ours))
    back = ref.RefTensor.new(idx_a, dtype).read_file(f_ours)
    assert back.indexes_equal(a)
    assert all(np.array_equal(x, y) for x, y in zip(back.blocks(), a.blocks()))
    assert np.array_equal(back.raw(), a.raw())


def test_contraction_result_files(ref, tmp_path):
    """A result of the reference's Contract (incl. a rank-0 scalar) survives the trip through our reader / writer."""
    # every rank-0 result of the case list (empty and non-empty scalars) plus a sample of the others, and the
    # reference's own full-trace fixtures (test_ten_ctrct.cc "2d_trace" / "3d_3axes": non-empty scalars)
    picked = [c for i, c in enumerate(CASES) if i % 5 == 0 or len(c[2][0]) == len(c[2][1]) == len(c[2][2][0])]
    picked += [(k, dt, (ia, ib, ax, da, db)) for (k, name, ia, ib, ax, da, db) in util.fixed_cases()
               if name in ("1d", "2d_trace", "3d_3axes") for dt in (np.float64, np.complex128)]
    n_scalar_full = 0
    for kind_name, dtype, (idx_a, idx_b, axes, div_a, div_b) in picked:
        a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, 77)
        c = ref.contract(a, b, axes)
        n_scalar_full += int(c.rank == 0 and c.raw().size == 1)
        f = tmp_path / "c.qlten"
        c.write_file(f)
        got = qlten_io.load(str(f), util.KINDS[kind_name], dtype)
        C = c.to_bst()
        assert got.indexes == C.indexes
        if C.rank == 0 and C.data.size == 0:      # "empty scalar": written as one zero element (blk_spar_data_ten.h:820-824)
            assert np.array_equal(got.data, np.zeros(1, dtype))
        else:
            assert np.array_equal(got.data, C.data)
        if C.rank:
            assert got.same_structure(C)
        assert qlten_io.dumps(C) == f.read_bytes()
    assert n_scalar_full >= 4, "the non-empty rank-0 case must be exercised"


def test_hash_known_answers():
    """Spot values of the hash functions the format embeds (computed by the reference: see the byte-exact test)."""
    assert qlten_io.qn_hash(tk.U1, (0,)) == 0
    assert qlten_io.qn_hash(tk.U1, (1,)) == 1 << 31
    assert qlten_io.qn_hash(tk.U1, (-1,)) == (1 << 64) - 1            # rotation of all ones
    s = tk.QNSector((2,), 3)
    assert qlten_io.sector_hash(tk.U1, s) == ((2 << 31) ^ 3)


def test_bench_loads_a_chain_from_files(tmp_path):
    """bench.py --tensors DIR: the five H_eff operands come back from disk ready to be chained."""
    import bench
    from tensortoolkit_b200 import workloads as wl
    rng = np.random.default_rng(3)
    ti = wl.heff_tensor_indexes(wl.hubbard_indexes(40))
    src = {n: tk.BlockSparseTensor(ix, np.float64).random((0, 0), rng) for n, ix in ti.items()}
    for n, t in src.items():
        qlten_io.save(t, str(tmp_path / (n + ".qlten")))
    got = bench.load_tensors(str(tmp_path), "fU1U1QN", "f64")
    for n in src:
        assert got[n].indexes == src[n].indexes and got[n].same_structure(src[n]) and np.array_equal(got[n].data, src[n].data)
    shells = dict(got)
    for lhs, rhs, axes, out in wl.HEFF_STEPS:           # the matcher accepts the loaded tensors
        m = tk.Match(shells[lhs], shells[rhs], axes)
        shells[out] = m.result_shell(np.float64)
        m.close()
    assert shells["out"].indexes == got["psi"].indexes


import glob
import os

from tests.golden import io as gio

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "file_*.qlten")))


def test_golden_files_present():
    assert len(FILES) >= 3


@pytest.mark.parametrize("path", FILES)
def test_golden_files(path):
    """Files written by the reference (tests/golden/make_golden.py) and the tensors they hold: read back exactly,
    re-written byte for byte -- without the reference library."""
    g = gio.load_case(path[:-len(".qlten")] + ".npz")
    want = g["A"]
    got = qlten_io.load(path, want.indexes[0].kind, want.dtype)
    assert got.indexes == want.indexes and got.same_structure(want) and np.array_equal(got.data, want.data)
    with open(path, "rb") as f:
        assert qlten_io.dumps(want) == f.read()
