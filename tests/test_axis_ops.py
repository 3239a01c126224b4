"""Matrix-free axis operations: dmrg::ApplyRank2ToAxisPreserveOrder and
dmrg::ApplyTwoRank2ToAxesPreserveOrder (tensor_manipulation/dmrg/axis_ops.h:2889-3125; the reference's own tests:
tests/test_tensor_manipulation/test_dmrg_axis_ops.cc -- two-op == two sequential single-op applications, boundary and
middle axes, block-sparse operators).

Host part: the numpy restatement and the library's output topology against the reference run here.  GPU part: values
against the reference through the Python API and the C++ adapter, the two-op form against two chained Contracts, and the
H_eff MPO steps expressed with the pre-contracted two-site operator against the four-step chain."""
import numpy as np
import pytest

import tensortoolkit_b200 as tk
from tensortoolkit_b200 import workloads as wl
from tensortoolkit_b200.tensor import IN, OUT, Index, QNSector
from oracle import contract_np as onp
from tests import util

TOL = 1e-12
BOSONIC = ["U1", "U1U1", "Z2"]


def axis_case(kind_name, rng, two, big_op=False):
    """State tensor of rank 2..5 with one / two target axes and block-sparse rank-2 operators (random divergence)."""
    kind = util.KINDS[kind_name]
    rank = int(rng.integers(2, 6))
    pool = [util.base_index(kind, rng, big=big_op) for _ in range(3)]
    idxs = [pool[int(rng.integers(3))] if rng.random() < 0.5 else pool[int(rng.integers(3))].inverse() for _ in range(rank)]
    axes = [int(a) for a in rng.choice(rank, size=2 if two else 1, replace=False)]
    zero = tuple([0] * kind.nvals)
    divs = [zero, (1,)] if kind.name == "Z2QN" else ([zero, (1,), (-1,)] if kind.nvals == 1 else [zero, (1, 1), (1, -1)])
    ops = []
    for ax in axes:
        out_ix = util.base_index(kind, rng, big=big_op)
        out_ix = out_ix if rng.random() < 0.5 else out_ix.inverse()
        ops.append(([idxs[ax].inverse(), out_ix], divs[int(rng.integers(len(divs)))], ax))
    return idxs, divs[int(rng.integers(len(divs)))], ops


def cases(n, seed, two, big_op=False):
    rng = np.random.default_rng(seed)
    out = []
    for kind_name in BOSONIC:
        for i in range(n):
            out.append((kind_name, np.float64 if i % 2 == 0 else np.complex128, axis_case(kind_name, rng, two, big_op)))
    return out


ONE = cases(8, 20261201, False)
TWO = cases(8, 20261202, True)
BIG = cases(2, 20261203, False, big_op=True) + cases(2, 20261204, True, big_op=True)


def build(ref, dtype, case, seed):
    idxs, div, ops = case
    ref.set_seed(seed)
    x = ref.RefTensor.new(idxs, dtype).random(div)
    rops = [(ref.RefTensor.new(oi, dtype).random(od), ax) for oi, od, ax in ops]
    return x, rops


@pytest.mark.parametrize("case", range(len(ONE) + len(TWO)))
def test_oracle_and_topology_vs_reference(ref, case):
    kind_name, dtype, cs = (ONE + TWO)[case]
    x, rops = build(ref, dtype, cs, 100 + case)
    if any(o.raw().size == 0 for o, _ in rops) or x.raw().size == 0:
        pytest.skip("empty operand (the reference requires non-default tensors)")
    want = ref.apply_rank2(x, rops).to_bst()
    X = x.to_bst(); ops = [(o.to_bst(), ax) for o, ax in rops]
    mine = onp.apply_rank2_axes_np(X, ops)
    assert mine.same_structure(want)
    if want.data.size:
        assert util.rel_fro(mine.data, want.data) <= TOL
    # the library's pairing / output topology (host only)
    plan = tk.AxisPlan(None, X, ops[0][0], ops[0][1], *(ops[1] if len(ops) > 1 else (None, -1)), host_only=True)
    shell = plan.result_shell()
    assert shell.same_structure(want)
    plan.close()


def test_precondition_errors():
    rng = np.random.default_rng(1)
    ix = Index(tk.U1, [QNSector((0,), 2), QNSector((1,), 3)], OUT)
    x = tk.BlockSparseTensor([ix, ix.inverse(), ix], np.float64).random((0,), rng)
    good = tk.BlockSparseTensor([ix.inverse(), ix], np.float64).random((0,), rng)
    bad = tk.BlockSparseTensor([ix, ix], np.float64).random((0,), rng)
    with pytest.raises(ValueError):
        tk.AxisPlan(None, x, bad, 0, host_only=True)                # op input index is not the inverse of the axis
    with pytest.raises(ValueError):
        tk.AxisPlan(None, x, good, 3, host_only=True)               # axis out of range
    with pytest.raises(ValueError):
        tk.AxisPlan(None, x, good, 0, good, 0, host_only=True)      # same axis twice
    fx = tk.BlockSparseTensor([Index(tk.fU1, [QNSector((0,), 2)], OUT)], np.float64)
    with pytest.raises(TypeError):
        tk.AxisPlan(None, fx, good, 0, host_only=True)              # bosonic only


# ---- GPU -----------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(ONE) + len(TWO) + len(BIG)))
def test_axis_ops_vs_reference(ref, ctx, case):
    kind_name, dtype, cs = (ONE + TWO + BIG)[case]
    x, rops = build(ref, dtype, cs, 100 + case)
    if any(o.raw().size == 0 for o, _ in rops) or x.raw().size == 0:
        pytest.skip("empty operand")
    want = ref.apply_rank2(x, rops)
    X = x.to_bst(); ops = [(o.to_bst(), ax) for o, ax in rops]
    if len(ops) == 1:
        got = tk.apply_rank2_to_axis_preserve_order(X, ops[0][0], ops[0][1], ctx)
    else:
        got = tk.apply_two_rank2_to_axes_preserve_order(X, ops[0][0], ops[0][1], ops[1][0], ops[1][1], ctx)
    util.assert_same_as_ref(got, want, TOL)
    # the C++ adapter on the reference's own tensors
    got2 = ref.b200_apply_rank2(x, rops, ctx.h)
    util.assert_same_as_ref(got2.to_bst(), want, TOL)
    if len(ops) == 2:
        # the reference's property test: the fused form equals two sequential single-axis applications
        seq = tk.apply_rank2_to_axis_preserve_order(tk.apply_rank2_to_axis_preserve_order(X, ops[0][0], ops[0][1], ctx), ops[1][0], ops[1][1], ctx)
        assert seq.same_structure(got) and (got.data.size == 0 or util.rel_fro(seq.data, got.data) <= TOL)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_site_operators_on_mps_tensor_d300(ref, ctx, dtype):
    """The DMRG use: two site operators on the physical legs of a two-site wave function psi[vb, ph, ph, vb] at D = 300
    (one launch, one pass), against the reference and against two chained Contract + Transpose calls."""
    ix = wl.u1_heisenberg_indexes(300)
    ti = wl.heff_tensor_indexes(ix)
    ref.set_seed(5)
    psi = ref.RefTensor.new(ti["psi"], dtype).random((0,))
    # site operators {ph IN, ph OUT}: Sz-like (div 0) on site 1, S+-like (div +2) on site 2
    o1 = ref.RefTensor.new([ix["ph_in"], ix["ph_out"]], dtype).random((0,))
    o2 = ref.RefTensor.new([ix["ph_in"], ix["ph_out"]], dtype).random((2,))
    want = ref.apply_rank2(psi, [(o1, 1), (o2, 2)])
    P, O1, O2 = psi.to_bst(), o1.to_bst(), o2.to_bst()
    plan = tk.AxisPlan(ctx, P, O1, 1, O2, 2)
    out = plan.result_shell()
    plan.execute_host(P.data, O1.data, O2.data, out.data)
    rd, wr = plan.bytes()
    assert wr == out.data.nbytes and rd >= P.data.nbytes // 2
    plan.close()
    util.assert_same_as_ref(out, want, TOL)
    c = tk.contract(tk.contract(P, O1, ([1], [0]), ctx), O2, ([1], [0]), ctx)      # (vb, vb, ph', ph'') -> back into place
    c = tk.transpose(c, [0, 2, 3, 1], ctx)
    assert c.same_structure(out) and util.rel_fro(c.data, out.data) <= TOL
