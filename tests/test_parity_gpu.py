"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle.

Bar (BASELINE.json north_star): bit-exact block structure / QN sectors / transposes, relative
Frobenius error <= 1e-12 for double and complex-double data.  The reference's own unit tests use
1e-13 abs/rel per element on tiny blocks (tests/testing_utility.h:19); Frobenius 1e-12 is the
north-star tolerance and is what every assertion below states.
"""
import glob
import os

import numpy as np
import pytest

import tensortoolkit_b200 as tk
from tensortoolkit_b200 import _lib, workloads as wl
from oracle import contract_np as onp
from tests import util
from tests.golden import io as gio

pytestmark = pytest.mark.gpu
TOL = 1e-12

CASES = util.case_list(n_per_kind=6)
BIG = util.case_list(n_per_kind=3, seed=99, big=True)
FIXED = util.fixed_cases()
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith(("contig_", "file_")))     # those belong to test_contiguous / test_qlten_io


@pytest.mark.parametrize("case", range(len(CASES)))
def test_random_cases_vs_reference(ref, ctx, case):
    kind_name, dtype, (idx_a, idx_b, axes, div_a, div_b) = CASES[case]
    a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, 1000 + case)
    c = tk.contract(a.to_bst(), b.to_bst(), axes, ctx)
    util.assert_same_as_ref(c, ref.contract(a, b, axes), TOL)


@pytest.mark.parametrize("case", range(len(BIG)))
def test_bigger_blocks_vs_reference(ref, ctx, case):
    kind_name, dtype, (idx_a, idx_b, axes, div_a, div_b) = BIG[case]
    a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, 2000 + case)
    c = tk.contract(a.to_bst(), b.to_bst(), axes, ctx)
    util.assert_same_as_ref(c, ref.contract(a, b, axes), TOL)


@pytest.mark.parametrize("case", range(len(FIXED)))
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_reference_fixture_shapes(ref, ctx, case, dtype):
    kind_name, name, idx_a, idx_b, axes, div_a, div_b = FIXED[case]
    a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, 7)
    c = tk.contract(a.to_bst(), b.to_bst(), axes, ctx)
    util.assert_same_as_ref(c, ref.contract(a, b, axes), TOL)


@pytest.mark.parametrize("case", range(0, len(CASES), 2))
def test_dropin_adapter_on_reference_tensors(ref, ctx, case):
    """qlten::b200::Contract(&A, &B, axes, &C) on the reference's own QLTensor objects."""
    kind_name, dtype, (idx_a, idx_b, axes, div_a, div_b) = CASES[case]
    a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, 3000 + case)
    want = ref.contract(a, b, axes)
    got = ref.b200_contract(a, b, axes, ctx.h)
    assert got.is_default() == want.is_default() or want.raw().size == 0
    util.assert_same_as_ref(got.to_bst(), want, TOL)


@pytest.mark.parametrize("path", GOLDEN)
def test_golden_vectors(ctx, path):
    g = gio.load_case(path)
    c = tk.contract(g["A"], g["B"], g["axes"], ctx)
    assert np.array_equal(c.blk_coors, g["C"].blk_coors) and np.array_equal(c.blk_offset, g["C"].blk_offset)
    if g["C"].data.size:
        assert util.rel_fro(c.data, g["C"].data) <= TOL


@pytest.mark.parametrize("case", range(len(CASES)))
def test_transpose_bit_exact(ref, ctx, case):
    """Whole-tensor transpose incl. fermionic reorder signs: pure data movement, bit-exact vs HPTT."""
    kind_name, dtype, (idx_a, _, _, div_a, _) = CASES[case]
    ref.set_seed(case)
    a = ref.RefTensor.new(idx_a, dtype).random(div_a)
    rng = np.random.default_rng(case)
    perm = [int(x) for x in rng.permutation(len(idx_a))]
    got = tk.transpose(a.to_bst(), perm, ctx)
    want = a.clone().transpose(perm)
    util.assert_same_as_ref(got, want, 0.0)
    if a.raw().size and perm != sorted(perm):
        got2 = a.clone().b200_transpose(perm, ctx.h)
        util.assert_same_as_ref(got2.to_bst(), want, 0.0)


def test_transpose_round_trip_and_shapes(ctx):
    """Ragged ranks/extents incl. size-1 axes, odd extents, rank 6; transpose then inverse == identity."""
    rng = np.random.default_rng(3)
    from tensortoolkit_b200.tensor import OUT, Index, QNSector
    for dtype in (np.float64, np.complex128):
        for shape in [(7,), (5, 3), (1, 9, 1), (33, 65, 3), (3, 1, 1, 130), (2, 3, 4, 5, 6), (3, 2, 1, 4, 2, 5), (257, 129)]:
            idxs = [Index(tk.U1, [QNSector((0,), d)], OUT) for d in shape]
            t = tk.BlockSparseTensor(idxs, dtype).random((0,), rng)
            for _ in range(3):
                perm = [int(x) for x in rng.permutation(len(shape))]
                out = tk.transpose(t, perm, ctx)
                want = np.transpose(t.block(0), perm)
                assert np.array_equal(out.block(0), want)
                inv = [perm.index(i) for i in range(len(perm))]
                back = tk.transpose(out, inv, ctx)
                assert np.array_equal(back.data, t.data)


def test_contract_1sector_sums_to_contract(ref, ctx):
    """test_ten_ctrct_1sct.cc:72-119: sum over sectors of Contract1Sector == Contract."""
    rng = np.random.default_rng(5)
    for kind_name, dtype in (("U1", np.float64), ("fU1U1", np.complex128)):
        idx_a, idx_b, axes, div_a, div_b = util.random_case(kind_name, rng, rank_a=3, rank_b=3, nctrct=1, big=True)
        a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, 11)
        A, B = a.to_bst(), b.to_bst()
        full = tk.contract(A, B, axes, ctx)
        dense = np.zeros_like(full.to_dense())
        free = [i for i in range(3) if i not in axes[0]][0]
        for s in range(idx_a[free].nsct):
            part = tk.contract_1sector(A, free, s, B, axes, ctx)
            util.assert_same_as_ref(part, ref.contract_1sector(a, free, s, b, axes), TOL)
            got = ref.b200_contract_1sector(a, free, s, b, axes, ctx.h)
            util.assert_same_as_ref(got.to_bst(), ref.contract_1sector(a, free, s, b, axes), TOL)
            dense += part.to_dense()
        assert util.rel_fro(dense, full.to_dense()) <= TOL


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("D", [64, 300])
def test_heff_apply_u1_vs_reference(ref, ctx, dtype, D):
    """Config 2/3 at reduced D: the four chained Contracts of a two-site H_eff apply."""
    ti = wl.heff_tensor_indexes(wl.u1_heisenberg_indexes(D))
    ref.set_seed(20260002)
    r = {name: ref.RefTensor.new(idxs, dtype).random((0,)) for name, idxs in ti.items()}
    t = {name: x.to_bst() for name, x in r.items()}
    for lhs, rhs, axes, out in wl.HEFF_STEPS:
        r[out] = ref.contract(r[lhs], r[rhs], axes)
        t[out] = tk.contract(t[lhs], t[rhs], axes, ctx)
        util.assert_same_as_ref(t[out], r[out], TOL)
    assert t["out"].indexes == t["psi"].indexes


def test_heff_apply_hubbard_fermionic_vs_reference(ref, ctx):
    """Config 4 at reduced D: fU1U1QN Grassmann tensors, many small sectors, f_ex_sign = -1 tasks."""
    ti = wl.heff_tensor_indexes(wl.hubbard_indexes(200))
    ref.set_seed(20260004)
    r = {name: ref.RefTensor.new(idxs, np.float64).random((0, 0)) for name, idxs in ti.items()}
    t = {name: x.to_bst() for name, x in r.items()}
    nneg = 0
    for lhs, rhs, axes, out in wl.HEFF_STEPS:
        m = tk.Match(t[lhs], t[rhs], axes)
        nneg += sum(1 for x in m.tasks() if x.sign < 0)
        r[out] = ref.contract(r[lhs], r[rhs], axes)
        t[out] = tk.contract(t[lhs], t[rhs], axes, ctx)
        util.assert_same_as_ref(t[out], r[out], TOL)
    assert nneg > 0


def test_dmma_path_equals_skinny_path(ctx):
    """Narrow pairs may run on either kernel; both must agree with the oracle."""
    rng = np.random.default_rng(8)
    ti = wl.heff_tensor_indexes(wl.u1_heisenberg_indexes(48))
    for dtype in (np.float64, np.complex128):
        t = {name: tk.BlockSparseTensor(idxs, dtype).random((0,), rng) for name, idxs in ti.items()}
        t1 = tk.contract(t["lenv"], t["psi"], ([0], [0]), ctx)
        m = tk.Match(t1, t["mpo1"], ([0, 2], [0, 1]))
        want = onp.contract_np(t1, t["mpo1"], ([0, 2], [0, 1]))
        for flags in (_lib.PLAN_DETERMINISTIC, _lib.PLAN_DETERMINISTIC | _lib.PLAN_NO_SKINNY):
            plan = tk.ContractionPlan(ctx, m, dtype, flags)
            st = plan.stats()
            assert (st.nrow_skinny > 0) == (flags == _lib.PLAN_DETERMINISTIC)
            c = m.result_shell(dtype)
            plan.execute_host(t1.data, t["mpo1"].data, c.data)
            assert util.rel_fro(c.data, want.data) <= TOL
            plan.close()


def test_deterministic_repeat(ctx):
    """Deterministic mode: two executions give bit-identical results (no atomics, fixed order)."""
    rng = np.random.default_rng(9)
    ti = wl.heff_tensor_indexes(wl.u1_heisenberg_indexes(200))
    t = {name: tk.BlockSparseTensor(idxs, np.complex128).random((0,), rng) for name, idxs in ti.items()}
    c1 = tk.contract(t["lenv"], t["psi"], ([0], [0]), ctx)
    c2 = tk.contract(t["lenv"], t["psi"], ([0], [0]), ctx)
    assert np.array_equal(c1.data, c2.data)


def test_permute_pass_on_device_buffers_that_are_only_8_byte_aligned(ctx):
    """The permute kernel's bulk-copy path (cp.async.bulk) needs 16-byte aligned base pointers; the plan only knows element
    offsets.  Double operands handed over at an odd element of a larger allocation must take the element path and still give
    the same numbers as the aligned call (which takes the bulk path: every run here is 16-byte aligned)."""
    import ctypes as C
    from tensortoolkit_b200.heff import DeviceBuffer
    rng = np.random.default_rng(7)
    n1, k, n2, m = 6, 40, 8, 24                      # B stored (n1, k, n2), contracted over k; runs of n2 = 8 doubles
    a = rng.random((m, k)); b = rng.random((n1, k, n2))
    want = (a @ np.transpose(b, (1, 0, 2)).reshape(k, n1 * n2)).reshape(-1)
    t = _lib.Task()
    t.a_ord = t.b_ord = t.c_ord = 0
    t.a_off = t.b_off = t.c_off = 0
    t.m, t.k, t.n, t.sign, t.first = m, k, n1 * n2, 1, 1
    tarr = (_lib.Task * 1)(t)
    ash = np.array([(m, k)], np.uint32); bsh = np.array([(n1, k, n2)], np.uint32)
    zero = np.zeros(1, np.uint64)
    h = C.c_void_p()
    _lib.check(_lib.lib.qlb200_plan_create_raw(
        ctx.h, _lib.F64, _lib.PLAN_DETERMINISTIC | 256, 2, (C.c_int32 * 2)(0, 1), 1, ash.ctypes.data_as(C.POINTER(C.c_uint32)),
        zero.ctypes.data_as(C.POINTER(C.c_uint64)), 3, (C.c_int32 * 3)(1, 0, 2), 1, bsh.ctypes.data_as(C.POINTER(C.c_uint32)),
        zero.ctypes.data_as(C.POINTER(C.c_uint64)), 1, tarr, m * n1 * n2, C.byref(h)), "plan_create_raw")     # 256 = QLB200_PLAN_NO_VIEW
    bufs = [DeviceBuffer(ctx, x.nbytes + 16) for x in (a, b)] + [DeviceBuffer(ctx, want.nbytes + 16)]
    for shift in (0, 8):
        for buf, x in zip(bufs, (a, b)):
            _lib.check(_lib.lib.qlb200_memcpy_h2d(ctx.h, C.c_void_p(buf.ptr + shift), x.ctypes.data, x.nbytes), "h2d")
        _lib.check(_lib.lib.qlb200_execute(ctx.h, h, C.c_void_p(bufs[0].ptr + shift), C.c_void_p(bufs[1].ptr + shift),
                                           C.c_void_p(bufs[2].ptr + shift), _lib.MEM_DEVICE), "execute")
        got = np.empty_like(want)
        _lib.check(_lib.lib.qlb200_memcpy_d2h(ctx.h, got.ctypes.data, C.c_void_p(bufs[2].ptr + shift), got.nbytes), "d2h")
        ctx.sync()
        assert util.rel_fro(got, want) <= TOL, shift
    _lib.lib.qlb200_plan_destroy(h)
    for buf in bufs:
        buf.free()


def test_ragged_raw_plan_vs_numpy(ctx):
    """Config 5 in miniature: descriptor table built directly (no shells), ragged m/k/n in [1, 300],
    several pairs per output block, rank-3 operand blocks with the config-5 permutations."""
    rng = np.random.default_rng(20260005)
    import ctypes as C
    for dtype, code in ((np.float64, _lib.F64), (np.complex128, _lib.C64)):
        nC, pairs = 40, 3
        a_shape, b_shape, a_off, b_off, tasks = [], [], [], [], []
        ao = bo = co = 0
        want = []
        A_chunks, B_chunks = [], []
        for c in range(nC):
            m = int(np.exp(rng.uniform(0, np.log(300)))); n = int(np.exp(rng.uniform(0, np.log(300))))
            m1 = max(d for d in range(1, int(m ** 0.5) + 1) if m % d == 0); m2 = m // m1
            n1 = max(d for d in range(1, int(n ** 0.5) + 1) if n % d == 0); n2 = n // n1
            acc = np.zeros((m, n), dtype)
            for p in range(pairs):
                k = int(np.exp(rng.uniform(0, np.log(300))))
                a = rng.random((k, m1, m2)).astype(dtype); b = rng.random((n1, k, n2)).astype(dtype)
                if dtype == np.complex128:
                    a = a + 1j * rng.random(a.shape); b = b + 1j * rng.random(b.shape)
                sign = -1 if rng.random() < 0.3 else 1
                acc += sign * (np.transpose(a, (1, 2, 0)).reshape(m, k) @ np.transpose(b, (1, 0, 2)).reshape(k, n))
                t = _lib.Task()
                t.a_ord, t.b_ord, t.c_ord = len(a_shape), len(b_shape), c
                t.a_off, t.b_off, t.c_off = ao, bo, co
                t.m, t.k, t.n, t.sign, t.first = m, k, n, sign, 1 if p == 0 else 0
                tasks.append(t)
                a_shape.append((k, m1, m2)); b_shape.append((n1, k, n2)); a_off.append(ao); b_off.append(bo)
                A_chunks.append(a.reshape(-1)); B_chunks.append(b.reshape(-1))
                ao += a.size; bo += b.size
            want.append(acc.reshape(-1)); co += m * n
        A = np.concatenate(A_chunks); B = np.concatenate(B_chunks); want = np.concatenate(want)
        Cbuf = np.empty(co, dtype)
        ash = np.array(a_shape, np.uint32); bsh = np.array(b_shape, np.uint32)
        aof = np.array(a_off, np.uint64); bof = np.array(b_off, np.uint64)
        tarr = (_lib.Task * len(tasks))(*tasks)
        h = C.c_void_p()
        _lib.check(_lib.lib.qlb200_plan_create_raw(
            ctx.h, code, _lib.PLAN_DETERMINISTIC, 3, (C.c_int32 * 3)(1, 2, 0), len(a_shape),
            ash.ctypes.data_as(C.POINTER(C.c_uint32)), aof.ctypes.data_as(C.POINTER(C.c_uint64)),
            3, (C.c_int32 * 3)(1, 0, 2), len(b_shape), bsh.ctypes.data_as(C.POINTER(C.c_uint32)),
            bof.ctypes.data_as(C.POINTER(C.c_uint64)), len(tasks), tarr, co, C.byref(h)), "plan_create_raw")
        _lib.check(_lib.lib.qlb200_execute(ctx.h, h, A.ctypes.data, B.ctypes.data, Cbuf.ctypes.data, _lib.MEM_HOST), "execute")
        _lib.lib.qlb200_plan_destroy(h)
        assert util.rel_fro(Cbuf, want) <= TOL


def test_partitioned_plans_cover_output(ctx):
    """Multi-GPU sharding on one device: the row slabs of all ranks tile C exactly and reproduce it."""
    rng = np.random.default_rng(12)
    ti = wl.heff_tensor_indexes(wl.u1_heisenberg_indexes(300))
    t = {name: tk.BlockSparseTensor(idxs, np.float64).random((0,), rng) for name, idxs in ti.items()}
    axes = ([0], [0])
    full = tk.contract(t["lenv"], t["psi"], axes, ctx)
    m = tk.Match(t["lenv"], t["psi"], axes)
    world = 4
    out = np.full(full.data.size, np.nan)
    covered = np.zeros(full.data.size, np.int32)
    for rank in range(world):
        plan = tk.ContractionPlan(ctx, m, np.float64)
        plan.partition(world, rank)
        off, ln = plan.c_ranges()
        for o, l in zip(off, ln):
            covered[int(o):int(o + l)] += 1
        plan.execute_host(t["lenv"].data, t["psi"].data, out)
        plan.close()
    assert np.all(covered == 1)
    assert np.array_equal(out, full.data)


def test_full_size_properties_heff_d4096_complex(ctx):
    """BASELINE headline size (U(1), D=4096, complex double): size-independent properties.
    Linearity of the H_eff apply in psi and agreement of the two GEMM kernels' block structure with
    psi's; the oracle cannot run this size in seconds, so no element-wise comparison here."""
    rng = np.random.default_rng(20260003)
    ti = wl.heff_tensor_indexes(wl.u1_heisenberg_indexes(4096))
    t = {name: tk.BlockSparseTensor(idxs, np.complex128).random((0,), rng) for name, idxs in ti.items()}

    def apply(psi):
        cur = dict(t, psi=psi)
        for lhs, rhs, axes, out in wl.HEFF_STEPS:
            cur[out] = tk.contract(cur[lhs], cur[rhs], axes, ctx)
        return cur["out"]

    psi2 = tk.BlockSparseTensor(ti["psi"], np.complex128).random((0,), rng)
    y1, y2 = apply(t["psi"]), apply(psi2)
    both = tk.BlockSparseTensor(ti["psi"], np.complex128)
    both.set_blocks(t["psi"].blk_coors, t["psi"].data + (0.5 - 2j) * psi2.data)
    y12 = apply(both)
    assert y1.indexes == t["psi"].indexes and np.array_equal(y1.blk_coors, t["psi"].blk_coors)
    assert util.rel_fro(y12.data, y1.data + (0.5 - 2j) * y2.data) <= TOL


@pytest.mark.parametrize("dtype", [np.complex128, np.float64])
@pytest.mark.parametrize("variant", ["ws", "ws_4m", "ws_stagger", "ws_4m_stagger", "ws_stream_k", "ws_4m_stream_k", "ws_permute_all",
                                     "ws_4m_permute_all"])
def test_gemm_kernel_variants(ref, ctx, variant, dtype):
    """The warp-specialised kernels (blocks read in place or through the permute kernel) against the
    reference on a fermionic chain with ragged K tails, ragged tile edges, -1 exchange signs (accumulator
    sign frames flipping inside a tile) and several pairs per block."""
    flags = {"ws": 0, "ws_permute_all": _lib.PLAN_PERMUTE_ALL,
             "ws_4m": _lib.PLAN_CPLX_4M, "ws_4m_permute_all": _lib.PLAN_CPLX_4M | _lib.PLAN_PERMUTE_ALL,
             "ws_stream_k": _lib.PLAN_STREAM_K, "ws_4m_stream_k": _lib.PLAN_CPLX_4M | _lib.PLAN_STREAM_K,
             "ws_stagger": _lib.PLAN_STAGGER_OUTPUT, "ws_4m_stagger": _lib.PLAN_CPLX_4M | _lib.PLAN_STAGGER_OUTPUT}[variant]
    flags |= _lib.PLAN_DETERMINISTIC | _lib.PLAN_NO_SKINNY
    ti = wl.heff_tensor_indexes(wl.hubbard_indexes(150))
    ref.set_seed(77)
    r = {name: ref.RefTensor.new(idxs, dtype).random((0, 0)) for name, idxs in ti.items()}
    t = {name: x.to_bst() for name, x in r.items()}
    for lhs, rhs, axes, out in wl.HEFF_STEPS:
        r[out] = ref.contract(r[lhs], r[rhs], axes)
        m = tk.Match(t[lhs], t[rhs], axes)
        plan = tk.ContractionPlan(ctx, m, dtype, flags)
        c = m.result_shell(dtype)
        plan.execute_host(t[lhs].data, t[rhs].data, c.data)
        plan.close(); m.close()
        util.assert_same_as_ref(c, r[out], TOL)
        t[out] = c


@pytest.mark.parametrize("dtype", [np.complex128, np.float64])
def test_transposed_operand_modes_vs_numpy(ctx, dtype):
    """Raw plans whose A blocks are stored k x m and whose B blocks are stored n x k (both read in
    place, transposed, by the GEMM producer), with ragged sizes around the tile edges; against numpy."""
    rng = np.random.default_rng(5)
    cplx = dtype == np.complex128
    sizes = [(1, 1, 1), (7, 5, 3), (33, 17, 129), (64, 40, 128), (65, 9, 131), (100, 70, 260), (31, 8, 127), (200, 33, 40)]
    a_shape, b_shape, a_off, b_off, tasks = [], [], [], [], []
    ao = bo = co = 0
    a_data, b_data, want = [], [], []
    for i, (m, k, n) in enumerate(sizes):
        a = rng.standard_normal((k, m)) + (1j * rng.standard_normal((k, m)) if cplx else 0)     # stored k x m
        b = rng.standard_normal((n, k)) + (1j * rng.standard_normal((n, k)) if cplx else 0)     # stored n x k
        a_shape += [k, m]; b_shape += [n, k]; a_off.append(ao); b_off.append(bo)
        tasks.append(dict(a_ord=i, b_ord=i, a_off=ao, b_off=bo, c_off=co, m=m, k=k, n=n, sign=-1 if i % 3 == 2 else 1, first=1))
        want.append((-1 if i % 3 == 2 else 1) * (a.T @ b.T))
        a_data.append(a.ravel()); b_data.append(b.ravel())
        ao += a.size; bo += b.size; co += m * n
    A = np.concatenate(a_data).astype(dtype); B = np.concatenate(b_data).astype(dtype)
    Cw = np.concatenate([w.ravel() for w in want]).astype(dtype)
    for flags in (_lib.PLAN_DETERMINISTIC | _lib.PLAN_NO_SKINNY, _lib.PLAN_DETERMINISTIC,
                  _lib.PLAN_DETERMINISTIC | _lib.PLAN_CPLX_4M | _lib.PLAN_NO_SKINNY,
                  _lib.PLAN_DETERMINISTIC | _lib.PLAN_PERMUTE_ALL | _lib.PLAN_NO_SKINNY):
        plan = tk.RawPlan(ctx, dtype, 2, [1, 0], a_shape, a_off, 2, [1, 0], b_shape, b_off, tasks, co, flags)
        Cg = np.zeros(co, dtype)
        plan.execute_host(A, B, Cg)
        plan.close()
        assert util.rel_fro(Cg, Cw) <= TOL


def test_chain_graph_replay_matches_eager(ctx):
    """qlb200_graph_*: the H_eff chain captured as one CUDA graph replays to bit-identical results, also after
    the input changed (the graph holds pointers, not data)."""
    from tensortoolkit_b200.heff import ContractionChain
    rng = np.random.default_rng(31)
    ti = wl.heff_tensor_indexes(wl.u1_heisenberg_indexes(120))
    t = {name: tk.BlockSparseTensor(idxs, np.complex128).random((0,), rng) for name, idxs in ti.items()}
    chain = ContractionChain(ctx, t, wl.HEFF_STEPS, np.complex128)
    chain.apply_device()
    eager = chain.result("out").data.copy()
    g = chain.capture()
    assert g.launches == chain.launches_per_apply and g.launches > 0
    g.launch()
    assert np.array_equal(chain.result("out").data, eager)
    psi2 = t["psi"].data * (0.25 - 1.5j)
    chain.upload("psi", psi2)
    g.launch()
    got = chain.result("out").data
    assert util.rel_fro(got, eager * (0.25 - 1.5j)) <= TOL
    g.close(); chain.close()


@pytest.mark.parametrize("dtype", [np.complex128, np.float64])
@pytest.mark.parametrize("extra", [0, _lib.PLAN_STAGGER_OUTPUT, _lib.PLAN_CPLX_4M, _lib.PLAN_NO_SPLIT_K, _lib.PLAN_STREAM_K,
                                   _lib.PLAN_STREAM_K | _lib.PLAN_CPLX_4M])
def test_split_k_long_contractions_vs_numpy(ctx, dtype, extra):
    """Few output blocks with very long k loops: the tiler cuts them along k (deterministic split-K: partial tiles
    + fix-up by the last arriver in slot order).  Ragged edges, several pairs per block, a -1 sign, identity and
    transposed operand storage; against numpy, and bit-identical when repeated."""
    rng = np.random.default_rng(77)
    cplx = dtype == np.complex128
    blocks = [(33, 100, [1500, 700]), (64, 200, [2100]), (20, 97, [900, 33, 1000])]
    a_shape, b_shape, a_off, b_off, tasks, want = [], [], [], [], [], []
    a_data, b_data = [], []
    ao = bo = co = 0
    for ci, (m, n, ks) in enumerate(blocks):
        acc = np.zeros((m, n), dtype)
        for pi, k in enumerate(ks):
            a = rng.standard_normal((m, k)) + (1j * rng.standard_normal((m, k)) if cplx else 0)
            b = rng.standard_normal((k, n)) + (1j * rng.standard_normal((k, n)) if cplx else 0)
            sign = -1 if (ci + pi) % 3 == 1 else 1
            acc += sign * (a @ b)
            tasks.append(dict(a_ord=len(a_off), b_ord=len(b_off), c_ord=ci, a_off=ao, b_off=bo, c_off=co, m=m, k=k, n=n,
                              sign=sign, first=1 if pi == 0 else 0))
            a_shape += [m, k]; b_shape += [k, n]; a_off.append(ao); b_off.append(bo)
            a_data.append(a.ravel()); b_data.append(b.ravel())
            ao += a.size; bo += b.size
        want.append(acc.ravel()); co += m * n
    A = np.concatenate(a_data).astype(dtype); B = np.concatenate(b_data).astype(dtype)
    Cw = np.concatenate(want).astype(dtype)
    flags = _lib.PLAN_DETERMINISTIC | _lib.PLAN_NO_SKINNY | extra
    plan = tk.RawPlan(ctx, dtype, 2, [0, 1], a_shape, a_off, 2, [0, 1], b_shape, b_off, tasks, co, flags)
    full_tiles = sum(-(-m // 32) * -(-n // 96) for m, n, _ in blocks)     # at least this many even with the widest tile
    if not (extra & _lib.PLAN_NO_SPLIT_K):
        assert plan.stats().ntile_dmma > 2 * full_tiles, "the k loops of this plan are expected to be cut"
    C1 = np.zeros(co, dtype); C2 = np.zeros(co, dtype)
    plan.execute_host(A, B, C1)
    plan.execute_host(A, B, C2)
    plan.close()
    assert util.rel_fro(C1, Cw) <= TOL
    assert np.array_equal(C1, C2)


@pytest.mark.parametrize("dtype", [np.complex128, np.float64])
def test_host_pipeline_matches_device_apply(ctx, dtype):
    """qlb200_hostpipe_*: psi streamed from host memory in chunks while the parts of step 1 run, the result streamed back in
    chunks while the last step computes -- bit-identical to the device-resident apply (same kernels, same summation order),
    also with different chunkings and when repeated with a new input."""
    from tensortoolkit_b200.heff import ContractionChain
    rng = np.random.default_rng(41)
    ti = wl.heff_tensor_indexes(wl.u1_heisenberg_indexes(260))
    t = {name: tk.BlockSparseTensor(idxs, dtype).random((0,), rng) for name, idxs in ti.items()}
    chain = ContractionChain(ctx, t, wl.HEFF_STEPS, dtype)
    chain.apply_device()
    want = chain.result("out").data.copy()
    for cum_in, cum_out in ((None, None), ((1.0,), (1.0,)), ((0.01, 0.5, 0.5, 0.99, 1.0), (0.3, 1.0)), ((0.5, 1.0), (0.1, 0.2, 0.3, 0.9, 1.0))):
        chain.make_host_pipe("psi", cum_in, cum_out)
        out = np.full(want.size, np.nan, dtype)
        n = chain.apply_host_pipelined(t["psi"].data, out)
        assert n > 0
        assert np.array_equal(out, want)
        psi2 = (t["psi"].data * 0.5).astype(dtype)
        chain.apply_host_pipelined(psi2, out)
        assert util.rel_fro(out, 0.5 * want) <= TOL
        tk._lib.lib.qlb200_hostpipe_destroy(chain._pipe); chain._pipe = None
    chain.close()


def test_plan_split_parts_cover_plan(ctx):
    """qlb200_plan_split: by operand arrival and by output range, the parts' output ranges tile C exactly once."""
    import ctypes as C
    rng = np.random.default_rng(43)
    ti = wl.heff_tensor_indexes(wl.hubbard_indexes(120))
    t = {name: tk.BlockSparseTensor(idxs, np.float64).random((0, 0), rng) for name, idxs in ti.items()}
    m = tk.Match(t["lenv"], t["psi"], ([0], [0]))
    plan = tk.ContractionPlan(ctx, m, np.float64)
    full = m.result_shell(np.float64)
    plan.execute_host(t["lenv"].data, t["psi"].data, full.data)
    for by in (_lib.SPLIT_BY_A, _lib.SPLIT_BY_B, _lib.SPLIT_BY_C):
        nparts = 5
        parts = (C.c_void_p * nparts)(); bounds = (C.c_uint64 * (nparts + 1))()
        cum = (C.c_double * nparts)(0.05, 0.3, 0.31, 0.8, 1.0)
        _lib.check(_lib.lib.qlb200_plan_split(plan.h, by, nparts, cum, parts, bounds), "plan_split")
        assert list(bounds) == sorted(bounds) and bounds[0] == 0
        cov = np.zeros(full.data.size, np.int32)
        out = np.full(full.data.size, np.nan)
        from tensortoolkit_b200.heff import DeviceBuffer
        dA, dB, dC = (DeviceBuffer(ctx, x.nbytes) for x in (t["lenv"].data, t["psi"].data, out))
        dA.upload(t["lenv"].data); dB.upload(t["psi"].data); dC.upload(out)
        for i in range(nparts):
            n = int(_lib.lib.qlb200_plan_c_range_count(parts[i]))
            off = np.zeros(n, np.uint64); ln = np.zeros(n, np.uint64)
            _lib.lib.qlb200_plan_c_ranges(parts[i], off.ctypes.data_as(C.POINTER(C.c_uint64)), ln.ctypes.data_as(C.POINTER(C.c_uint64)))
            for o, l in zip(off, ln):
                cov[int(o):int(o + l)] += 1
                if by == _lib.SPLIT_BY_C:
                    assert bounds[i] <= o and o + l <= bounds[i + 1]
            _lib.check(_lib.lib.qlb200_execute_gemm(ctx.h, parts[i], dA.ptr, dB.ptr, dC.ptr), "execute_gemm")
            _lib.lib.qlb200_plan_destroy(parts[i])
        dC.download(out); ctx.sync()
        for b in (dA, dB, dC):
            b.free()
        assert np.all(cov == 1)
        assert np.array_equal(out, full.data)
    plan.close(); m.close()
