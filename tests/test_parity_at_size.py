"""Parity AT SIZE: the BASELINE configurations themselves (not reduced twins) against the reference run on the same
inputs -- block topology equal, relative Frobenius error <= 1e-12 (north_star), permuted operands bit-exact.

  * headline: U(1) two-site H_eff apply, D = 4096, complex double (BASELINE configs[2]) -- every step of the chain;
  * fermionic Hubbard chain at D = 2048 (configs[3] at a quarter of its bond dimension: 58 sectors, f_ex_sign = -1 tasks);
  * ragged stress test (configs[4]): the first 1000 pairs of the SAME descriptor table through the reference's executor
    loop (hp_numeric::TensorTranspose + hp_numeric::MatMultiply, global_operations.h:919-982), permuted blocks bit-exact;
  * adversarial complex inputs for the 3M (Karatsuba) complex product against the reference's ZGEMM, with the
    componentwise behaviour that decides when a plan must use QLB200_PLAN_CPLX_4M.
The reference CPU runs take a few seconds each on the GPU box's host cores.
"""
import zlib

import numpy as np
import pytest

import tensortoolkit_b200 as tk
from tensortoolkit_b200 import _lib, workloads as wl
from tensortoolkit_b200.heff import ContractionChain
from tests import util

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _chain_vs_reference(ref, ctx, ti, dtype, div, seed):
    ref.set_seed(seed)
    r = {name: ref.RefTensor.new(idxs, dtype).random(div) for name, idxs in ti.items()}
    t = {name: x.to_bst() for name, x in r.items()}
    chain = ContractionChain(ctx, t, wl.HEFF_STEPS, dtype)
    chain.apply_device()
    ctx.sync()
    errs = {}
    for lhs, rhs, axes, out in wl.HEFF_STEPS:
        r[out] = ref.contract(r[lhs], r[rhs], axes)
        got = chain.result(out)
        util.assert_same_as_ref(got, r[out], TOL)           # indexes, block keys / coordinates / shapes / offsets, values
        errs[out] = util.rel_fro(got.data, r[out].raw())
        for name in (lhs, rhs):                             # free reference intermediates as soon as possible
            if name in ("t1", "t2", "t3"):
                r[name].free()
    chain.close()
    return errs


def test_headline_heff_apply_d4096_complex_vs_reference(ref, ctx):
    """BASELINE configs[2] at full size: all four chained Contracts against qlten::Contract on the host."""
    ti = wl.heff_tensor_indexes(wl.u1_heisenberg_indexes(4096))
    errs = _chain_vs_reference(ref, ctx, ti, np.complex128, (0,), 20260003)
    print("headline D=4096 complex, rel. Frobenius error per step:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) <= TOL


def test_headline_heff_apply_d4096_double_vs_reference(ref, ctx):
    """The real-double kernels on the headline block structure (config 2's structure at D = 4096)."""
    ti = wl.heff_tensor_indexes(wl.u1_heisenberg_indexes(4096))
    errs = _chain_vs_reference(ref, ctx, ti, np.float64, (0,), 20260002)
    assert max(errs.values()) <= TOL


def test_hubbard_heff_apply_d2048_vs_reference(ref, ctx):
    """BASELINE configs[3] structure (fU1U1QN Grassmann tensors, many small sectors) at D = 2048, double."""
    ti = wl.heff_tensor_indexes(wl.hubbard_indexes(2048))
    errs = _chain_vs_reference(ref, ctx, ti, np.float64, (0, 0), 20260004)
    print("Hubbard D=2048, rel. Frobenius error per step:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) <= TOL


def _ragged_slice(ntask):
    """The first `ntask` pairs of the config-5 table, offsets re-based to a compact slice."""
    tb = wl.ragged_tables()
    sub = tb["tasks"][:ntask].copy()
    a_off, b_off = np.zeros(ntask, np.uint64), np.zeros(ntask, np.uint64)
    ao = bo = co = 0
    last_c, c_base = None, {}
    for i, t in enumerate(sub):
        m, k, n = int(t["m"]), int(t["k"]), int(t["n"])
        a_off[i], b_off[i] = ao, bo
        if int(t["c_ord"]) not in c_base:
            c_base[int(t["c_ord"])] = co
            co += m * n
        t["a_ord"] = t["b_ord"] = t["a_blk_idx"] = t["b_blk_idx"] = i
        t["a_off"], t["b_off"], t["c_off"] = ao, bo, c_base[int(t["c_ord"])]
        ao += m * k; bo += k * n
    return dict(tasks=sub, a_shape=tb["a_shape"][:ntask], b_shape=tb["b_shape"][:ntask], a_off=a_off, b_off=b_off,
                a_elems=ao, b_elems=bo, c_elems=co)


@pytest.mark.parametrize("flags", [0, _lib.PLAN_NO_VIEW, _lib.PLAN_PERMUTE_ALL], ids=["in_place_and_views", "in_place_no_views", "permute_all"])
def test_ragged_slice_vs_reference_executor_loop(ref, ctx, flags):
    """1000 pairs of the ragged table: A stored (k, m1, m2) perm {1,2,0}, B stored (n1, k, n2) perm {1,0,2}.
    Values against HPTT + OpenBLAS; every block that goes through the permute kernel bit-exact against HPTT."""
    s = _ragged_slice(1000)
    rng = np.random.Generator(np.random.MT19937(20260005))
    A = rng.random(s["a_elems"]); B = rng.random(s["b_elems"])
    want, sec, At, Bt = ref.raw_contract(np.float64, 3, [1, 2, 0], s["a_shape"], s["a_off"], 3, [1, 0, 2], s["b_shape"], s["b_off"],
                                         s["tasks"], A, B, s["c_elems"], keep_permuted=True)
    plan = tk.RawPlan(ctx, np.float64, 3, [1, 2, 0], s["a_shape"], s["a_off"], 3, [1, 0, 2], s["b_shape"], s["b_off"], s["tasks"],
                      s["c_elems"], _lib.PLAN_DETERMINISTIC | flags)
    got = np.zeros(s["c_elems"])
    plan.execute_host(A, B, got)
    st = plan.stats()
    # permuted operands, bit for bit
    checked = 0
    for which, ref_t, offs, shapes in ((0, At, s["a_off"], s["a_shape"]), (1, Bt, s["b_off"], s["b_shape"])):
        elems = st.permute_elems_a if which == 0 else st.permute_elems_b
        if elems == 0:
            continue
        for b in range(len(offs)):
            wo = plan.operand_block(which, b)
            if wo is None:
                continue
            size = int(np.prod(shapes[b].astype(np.int64)))
            mine = plan.read_workspace(which, wo, size)
            assert np.array_equal(mine, ref_t[int(offs[b]):int(offs[b]) + size]), f"operand {which} block {b}: permuted copy differs from HPTT"
            checked += 1
    if flags & _lib.PLAN_PERMUTE_ALL:
        assert checked == 2 * len(s["tasks"])
    elif flags & _lib.PLAN_NO_VIEW:
        assert checked > len(s["tasks"]) // 2      # most (n1, k, n2) blocks of B go through the permute kernel
    else:
        # default: A blocks are 2-D transpositions and B blocks strided views, both read in place by the GEMM producers;
        # only B blocks whose inner run is 2 or 3 elements long are permuted
        assert st.permute_elems_a == 0 and st.permute_elems_b < 0.01 * s["b_elems"]
    plan.close()
    # values: whole slice and worst output block
    assert util.rel_fro(got, want) <= TOL
    worst = 0.0
    seen = set()
    for t in s["tasks"]:
        c0 = int(t["c_off"])
        if c0 in seen:
            continue
        seen.add(c0)
        n = int(t["m"]) * int(t["n"])
        worst = max(worst, util.rel_fro(got[c0:c0 + n], want[c0:c0 + n]))
    print(f"ragged slice: {len(s['tasks'])} pairs, {checked} permuted blocks bit-exact, worst block rel err {worst:.2e}, reference loop {sec:.2f} s")
    assert worst <= TOL


def _complex_blocks(rng, make):
    """A few output blocks with 2 pairs each, sizes around and beyond the tile edges; `make(shape)` draws the data."""
    sizes = [(96, 256, 192), (33, 517, 100), (200, 64, 97), (64, 1024, 96), (17, 40, 7)]
    tasks, a_shape, b_shape, a_off, b_off, ad, bd = [], [], [], [], [], [], []
    ao = bo = co = 0
    for ci, (m, k, n) in enumerate(sizes):
        for p in range(2):
            kk = k if p == 0 else max(4, k // 3)
            a, b = make((m, kk)), make((kk, n))
            tasks.append(dict(a_ord=len(a_off), b_ord=len(b_off), c_ord=ci, a_off=ao, b_off=bo, c_off=co, m=m, k=kk, n=n,
                              sign=-1 if (ci + p) % 3 == 0 else 1, first=1 if p == 0 else 0))
            a_shape.append((m, kk)); b_shape.append((kk, n)); a_off.append(ao); b_off.append(bo)
            ad.append(a.ravel()); bd.append(b.ravel()); ao += a.size; bo += b.size
        co += m * n
    return tasks, a_shape, b_shape, a_off, b_off, np.concatenate(ad), np.concatenate(bd), co, sizes


ADVERSARIAL = {
    # |Im| ~ 1e-8 |Re|: the imaginary part of the 3M product is a difference of three O(1) sums
    "tiny_imag": lambda rng: (lambda sh: rng.random(sh) + 1e-8j * rng.random(sh)),
    "tiny_real": lambda rng: (lambda sh: 1e-8 * rng.random(sh) + 1j * rng.random(sh)),
    # magnitudes spread over 1e-6 .. 1e+6, row by row and element by element
    "wide_range": lambda rng: (lambda sh: (rng.standard_normal(sh) + 1j * rng.standard_normal(sh)) * 10.0 ** rng.uniform(-6, 6, sh)),
    "wide_rows": lambda rng: (lambda sh: (rng.random(sh) + 1j * rng.random(sh)) * 10.0 ** rng.uniform(-6, 6, (sh[0], 1))),
    # sign-mixed data: cancellation inside every dot product
    "sign_mixed": lambda rng: (lambda sh: rng.standard_normal(sh) + 1j * rng.standard_normal(sh)),
    "re_im_anticorrelated": lambda rng: (lambda sh: (lambda x: x - 1j * x * (1 + 1e-9 * rng.standard_normal(sh)))(rng.standard_normal(sh))),
}


@pytest.mark.parametrize("kind", sorted(ADVERSARIAL))
def test_complex_3m_adversarial_inputs_vs_reference_zgemm(ref, ctx, kind):
    """3M (default) and 4M complex products against the reference's ZGEMM on inputs chosen to hurt 3M.

    Contract: both forms satisfy the NORMWISE bound |C^ - C|_F <= c k u |A|_F |B|_F, which is what the
    parity bar (relative Frobenius <= 1e-12 per tensor, here also per output block) measures.  3M does NOT give
    componentwise accuracy of the smaller of (Re, Im) when it is more than ~1e4 below the larger one: a caller who needs
    that must plan with QLB200_PLAN_CPLX_4M (or export QLB200_COMPLEX_PRODUCT=4m), which this test shows to be accurate
    component by component."""
    rng = np.random.default_rng(zlib.crc32(kind.encode()))
    make = ADVERSARIAL[kind](rng)
    tasks, a_shape, b_shape, a_off, b_off, A, B, c_elems, sizes = _complex_blocks(rng, make)
    want, _ = ref.raw_contract(np.complex128, 2, None, a_shape, a_off, 2, None, b_shape, b_off, tasks, A, B, c_elems)
    out = {}
    for name, fl in (("3m", 0), ("4m", _lib.PLAN_CPLX_4M)):
        plan = tk.RawPlan(ctx, np.complex128, 2, [0, 1], a_shape, a_off, 2, [0, 1], b_shape, b_off, tasks, c_elems,
                          _lib.PLAN_DETERMINISTIC | _lib.PLAN_NO_SKINNY | fl)
        got = np.zeros(c_elems, np.complex128)
        plan.execute_host(A, B, got)
        plan.close()
        out[name] = got
    report = {}
    for name, got in out.items():
        worst, co = 0.0, 0
        for (m, k, n) in sizes:
            worst = max(worst, util.rel_fro(got[co:co + m * n], want[co:co + m * n]))
            co += m * n
        small = np.minimum(np.abs(want.real), np.abs(want.imag))
        part = np.where(np.abs(want.real) < np.abs(want.imag), np.abs(got.real - want.real), np.abs(got.imag - want.imag))
        report[name] = dict(normwise=util.rel_fro(got, want), worst_block=worst, small_part=float(np.linalg.norm(part) / max(np.linalg.norm(small), 1e-300)))
    print(kind, {k: {a: f"{b:.1e}" for a, b in v.items()} for k, v in report.items()})
    for name in ("3m", "4m"):
        assert report[name]["normwise"] <= TOL and report[name]["worst_block"] <= TOL
    if kind in ("tiny_imag", "tiny_real"):
        # the documented limit of 3M: the small part is only accurate relative to the LARGE part ...
        assert report["3m"]["small_part"] <= 1e-6
        # ... and 4M is the fallback that keeps it accurate in its own right
        assert report["4m"]["small_part"] <= 1e-11
