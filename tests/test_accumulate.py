"""Accumulate forms of the contiguous-axes contraction (second half):
qlten::ContractTailHeadContiguousAccumulate / TryContractTailHeadContiguousAccumulate
(tensor_manipulation/contract_contiguous_axes.h:954-1041), the reference's tests for them:
tests/test_tensor_manipulation/test_ten_ctrct.cc:1083-1467 (identical topology, superset output, missing blocks ->
expansion, Try... probe, default output, beta rules, the QN<U1,U1> case).

Host part (no GPU): the output topology / first-task-beta table and the ContiguousContractStats counters computed by the
library (qlb200_accum_*) against the reference run here, and the numpy restatement against the reference's values.
GPU part: values through the Python API and through the C++ drop-in adapter against the reference.
"""
import numpy as np
import pytest

import tensortoolkit_b200 as tk
from tensortoolkit_b200 import _lib
from tensortoolkit_b200.contract import (AccumulateLayoutMismatch, contract_tail_head_contiguous_accumulate as accumulate,
                                         try_contract_tail_head_contiguous_accumulate as try_accumulate)
from oracle import contract_np as onp
from tests import util
from tests.test_contiguous import contiguous_case

TOL = 1e-12
SCALARS = {np.float64: [(1.0, 0.0), (0.7, 1.0), (-1.3, 0.4), (2.0, 0.0)],
           np.complex128: [(1.0, 0.0), (0.7 - 0.2j, 1.0), (-1.3 + 0.5j, 0.4 - 1.1j), (2.0j, 0.0)]}


def cases(n_per_kind, seed):
    rng = np.random.default_rng(seed)
    out = []
    for kind_name in util.KINDS:
        for i in range(n_per_kind):
            out.append((kind_name, np.float64 if i % 2 == 0 else np.complex128, contiguous_case(kind_name, rng, big=(i % 3 == 2))))
    return out


CASES = cases(6, 20261101)


def subset_of(c, rng, mode):
    """An 'existing output' derived from a full contraction result: identical / superset / subset block topology."""
    t = tk.BlockSparseTensor(c.indexes, c.dtype)
    if c.rank == 0:
        t.data = (rng.random(c.data.size) + (1j * rng.random(c.data.size) if c.dtype == np.complex128 else 0)).astype(c.dtype)
        return t
    if mode == "identical":
        coors = c.blk_coors
    elif mode == "superset":      # every block of the result's divergence class plus the result's own
        allowed = c.div_blocks(_div_of(c)) if c.nblk else np.zeros((0, c.rank), np.uint32)
        coors = np.unique(np.concatenate([c.blk_coors, allowed]), axis=0) if len(allowed) else c.blk_coors
    else:                         # "subset": drop about half of the blocks -> the accumulate must expand the topology
        keep = rng.random(c.nblk) < 0.5
        coors = c.blk_coors[keep]
    if len(coors):
        t.set_blocks(coors)
        n = t.data.size
        t.data = (rng.random(n) + (1j * rng.random(n) if c.dtype == np.complex128 else 0)).astype(c.dtype)
    return t


def _div_of(c):
    """Divergence of a non-empty tensor (CalcDiv): flow of its first block."""
    kind = c.kind
    tot = np.zeros(kind.nvals, np.int64)
    for i, ix in enumerate(c.indexes):
        tot += np.array(ix.sectors[int(c.blk_coors[0, i])].qn, np.int64) * ix.dir
    return tuple(int(v) for v in kind.norm(tuple(tot)))


def make_inputs(ref, kind_name, dtype, case, seed):
    idx_a, idx_b, (a_start, b_start, size), div_a, div_b = case
    a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, seed)
    return a, b, a_start, b_start, size


def run_reference(ref, a, b, a_start, b_start, size, alpha, beta, c_bst, try_only=False):
    c = ref.RefTensor.from_bst(c_bst, a.indexes[0].kind) if c_bst is not None else ref.default_like(a)
    ok, st = ref.contract_accumulate(a, b, a_start, b_start, size, alpha, beta, c, try_only)
    return ok, st, c


@pytest.mark.parametrize("mode", ["default", "identical", "superset", "subset"])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_layout_stats_and_oracle_vs_reference(ref, case, mode):
    """No GPU: resulting block topology, old/new offsets, touched flags, stats counters and the numpy restatement."""
    kind_name, dtype, cs = CASES[case]
    a, b, a_start, b_start, size = make_inputs(ref, kind_name, dtype, cs, 500 + case)
    A, B = a.to_bst(), b.to_bst()
    rng = np.random.default_rng(case)
    full = onp.contract_contiguous_np(A, B, a_start, b_start, size)
    for alpha, beta in SCALARS[dtype]:
        if mode == "default":
            c0 = None
            beta = 0.0
        else:
            if full.rank and full.nblk == 0:
                return
            c0 = subset_of(full, rng, mode)
        ok, rst, rc = run_reference(ref, a, b, a_start, b_start, size, alpha, beta, c0)
        want = rc.to_bst() if not rc.is_default() else None
        # numpy restatement == reference
        mine = onp.contract_accumulate_np(A, B, a_start, b_start, size, alpha, beta, c0)
        if want is None:
            assert mine.data.size == 0
        else:
            assert mine.same_structure(want)
            if want.data.size:
                assert util.rel_fro(mine.data, want.data) <= TOL
        # library layout + counters == reference
        m = tk.Match(A, B, None, contiguous=(a_start, b_start, size))
        import ctypes as C
        acc = C.c_void_p()
        al = (C.c_double * 2)(complex(alpha).real, complex(alpha).imag); be = (C.c_double * 2)(complex(beta).real, complex(beta).imag)
        sh = c0.shell() if c0 is not None else None
        _lib.check(_lib.lib.qlb200_accum_create(m.h, sh.ptr() if sh is not None else None, int(c0 is not None and c0.data.size > 0), 1,
                                                0 if dtype == np.float64 else 1, al, be, C.byref(acc)), "accum_create")
        st = _lib.AccumStats()
        _lib.check(_lib.lib.qlb200_accum_get_stats(acc, C.byref(st)), "stats")
        assert st.as_dict() == rst, f"alpha={alpha} beta={beta} mode={mode}"
        n = int(_lib.lib.qlb200_accum_nblk(acc))
        if want is not None and want.rank:
            assert n == want.nblk
            coors = np.zeros((n, want.rank), np.uint32); off = np.zeros(n, np.uint64); old = np.zeros(n, np.uint64); tch = np.zeros(n, np.uint8)
            _lib.check(_lib.lib.qlb200_accum_blocks(acc, None, coors.ctypes.data_as(C.POINTER(C.c_uint32)), None,
                                                    off.ctypes.data_as(C.POINTER(C.c_uint64)), old.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                    tch.ctypes.data_as(C.POINTER(C.c_uint8))), "blocks")
            assert np.array_equal(coors, want.blk_coors) and np.array_equal(off, want.blk_offset)
            assert int(_lib.lib.qlb200_accum_elems(acc)) == want.data.size
            req = {tuple(int(x) for x in r) for r in full.blk_coors}
            oldk = {tuple(int(x) for x in c0.blk_coors[i]): int(c0.blk_offset[i]) for i in range(c0.nblk)} if c0 is not None else {}
            for i in range(n):
                k = tuple(int(x) for x in coors[i])
                assert bool(tch[i]) == (k in req)
                assert (int(old[i]) if old[i] != np.uint64(2 ** 64 - 1) else None) == oldk.get(k)
        _lib.lib.qlb200_accum_destroy(acc)
        m.close()


def test_try_probe_and_argument_errors(ref):
    """Try... returns False (nothing changed) when blocks are missing; beta != 0 on a default output and index mismatch."""
    rng = np.random.default_rng(3)
    for kind_name, dtype, cs in CASES:          # first case whose result has several blocks
        a, b, a_start, b_start, size = make_inputs(ref, kind_name, dtype, cs, 77)
        A, B = a.to_bst(), b.to_bst()
        full = onp.contract_contiguous_np(A, B, a_start, b_start, size)
        if full.rank and full.nblk >= 2:
            break
    assert full.nblk >= 2
    sub = subset_of(full, np.random.default_rng(1), "subset")
    if sub.nblk == full.nblk:
        sub.set_blocks(full.blk_coors[:1]); sub.data[...] = 1.0
    ok, rst, rc = run_reference(ref, a, b, a_start, b_start, size, 1.0, 1.0, sub, try_only=True)
    assert ok is False
    with pytest.raises(onp.AccumulateLayoutMismatchNp):
        onp.contract_accumulate_np(A, B, a_start, b_start, size, 1.0, 1.0, sub, allow_expand=False)
    # library: same verdicts without touching a GPU (the layout step is host-only)
    import ctypes as C
    m = tk.Match(A, B, None, contiguous=(a_start, b_start, size))
    acc = C.c_void_p()
    one = (C.c_double * 2)(1.0, 0.0)
    rc_ = _lib.lib.qlb200_accum_create(m.h, sub.shell().ptr(), 1, 0, 0 if dtype == np.float64 else 1, one, one, C.byref(acc))
    assert rc_ == _lib.ERR_LAYOUT
    rc_ = _lib.lib.qlb200_accum_create(m.h, None, 0, 1, 0 if dtype == np.float64 else 1, one, one, C.byref(acc))
    assert rc_ == _lib.ERR_ARG        # default output needs beta == 0
    m.close()


# ---- GPU ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["default", "identical", "superset", "subset"])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_accumulate_vs_reference(ref, ctx, case, mode):
    kind_name, dtype, cs = CASES[case]
    a, b, a_start, b_start, size = make_inputs(ref, kind_name, dtype, cs, 500 + case)
    A, B = a.to_bst(), b.to_bst()
    rng = np.random.default_rng(case)
    full = onp.contract_contiguous_np(A, B, a_start, b_start, size)
    for alpha, beta in SCALARS[dtype]:
        if mode == "default":
            c0, beta = None, 0.0
        else:
            if full.rank and full.nblk == 0:
                return
            c0 = subset_of(full, rng, mode)
        ok, rst, rc = run_reference(ref, a, b, a_start, b_start, size, alpha, beta, c0)
        stats = {}
        got = accumulate(A, B, a_start, b_start, size, alpha, beta, c0, ctx, stats)
        assert stats == rst
        if rc.is_default():
            assert got.data.size == 0
            continue
        util.assert_same_as_ref(got, rc, TOL)
        # the no-expansion probe agrees with the reference's verdict
        ok_ref, _, rc2 = run_reference(ref, a, b, a_start, b_start, size, alpha, beta, c0, try_only=True)
        ok_me, got2 = try_accumulate(A, B, a_start, b_start, size, alpha, beta, c0, ctx)
        assert ok_me == ok_ref
        if ok_me:
            util.assert_same_as_ref(got2, rc2, TOL)
        else:
            assert got2 is c0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["default", "identical", "superset", "subset"])
@pytest.mark.parametrize("case", range(0, len(CASES), 2))
def test_dropin_adapter_accumulate(ref, ctx, case, mode):
    """qlten::b200::(Try)ContractTailHeadContiguousAccumulate on the reference's own QLTensor objects."""
    kind_name, dtype, cs = CASES[case]
    a, b, a_start, b_start, size = make_inputs(ref, kind_name, dtype, cs, 900 + case)
    A, B = a.to_bst(), b.to_bst()
    rng = np.random.default_rng(case)
    full = onp.contract_contiguous_np(A, B, a_start, b_start, size)
    alpha, beta = SCALARS[dtype][2]
    if mode == "default":
        c0, beta = None, 0.0
    else:
        if full.rank and full.nblk == 0:
            return
        c0 = subset_of(full, rng, mode)
    for try_only in (False, True):
        ok, rst, rc = run_reference(ref, a, b, a_start, b_start, size, alpha, beta, c0, try_only)
        mine = ref.RefTensor.from_bst(c0, a.indexes[0].kind) if c0 is not None else ref.default_like(a)
        ok2, st2 = ref.b200_contract_accumulate(a, b, a_start, b_start, size, alpha, beta, mine, try_only, ctx.h)
        assert ok2 == ok and st2 == rst
        assert mine.is_default() == rc.is_default()
        if not rc.is_default():
            util.assert_same_as_ref(mine.to_bst(), rc, TOL if ok else 0.0)


@pytest.mark.gpu
def test_accumulate_large_blocks_split_k_and_skinny(ctx):
    """The accumulate epilogue of every kernel family: DMMA tiles with and without split-K, narrow pairs, complex and real,
    against numpy (beta * C + alpha * A B) on a raw-descriptor-free path: H_eff step shapes at D = 300."""
    from tensortoolkit_b200 import workloads as wl
    rng = np.random.default_rng(4)
    ti = wl.heff_tensor_indexes(wl.u1_heisenberg_indexes(300))
    for dtype, alpha, beta in ((np.float64, -0.75, 0.5), (np.complex128, 0.3 - 1.2j, -0.4 + 0.9j)):
        t = {name: tk.BlockSparseTensor(idxs, dtype).random((0,), rng) for name, idxs in ti.items()}
        # step 1 (lenv x psi: big GEMMs, contracted axes lenv[0] / psi[0]) and step 2 shape (narrow pairs)
        t1 = tk.contract_contiguous_axes(t["lenv"], t["psi"], 0, 0, 1, ctx)
        want1 = onp.contract_accumulate_np(t["lenv"], t["psi"], 0, 0, 1, alpha, beta, t1)
        got1 = accumulate(t["lenv"], t["psi"], 0, 0, 1, alpha, beta, t1, ctx)
        assert got1.same_structure(want1) and util.rel_fro(got1.data, want1.data) <= TOL
        mp = tk.BlockSparseTensor([ti["mpo1"][1], ti["mpo1"][0], ti["mpo1"][2], ti["mpo1"][3]], dtype).random((0,), rng)
        # psi[vb, ph, ph, vb] x W'[ph IN, ...]: contract psi's axis 1 with W' axis 0 -> narrow-pair kernel
        t2 = tk.contract_contiguous_axes(t["psi"], mp, 1, 0, 1, ctx)
        want2 = onp.contract_accumulate_np(t["psi"], mp, 1, 0, 1, alpha, beta, t2)
        got2 = accumulate(t["psi"], mp, 1, 0, 1, alpha, beta, t2, ctx)
        assert got2.same_structure(want2) and util.rel_fro(got2.data, want2.data) <= TOL
