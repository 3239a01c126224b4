#!/usr/bin/env python
"""Small cases through every kernel family once, for compute-sanitizer (memcheck / racecheck / synccheck):
warp-specialised complex GEMM (3M, 4M, split-K + fix-up, staggered output, peer-store epilogue), real GEMM, narrow-pair
kernel, batched permute (all modes), whole-tensor transpose, accumulate epilogues + scale-copy, axis apply, host pipeline.
Results are checked against the numpy restatement, so a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensortoolkit_b200 as tk
from tensortoolkit_b200 import _lib, workloads as wl
from tensortoolkit_b200.heff import ContractionChain
from oracle import contract_np as onp

ctx = tk.Context(0)
rng = np.random.default_rng(1)


def rel(x, y):
    return float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300))


for name, ixf, div in (("u1", wl.u1_heisenberg_indexes, (0,)), ("hubbard", wl.hubbard_indexes, (0, 0))):
    ti = wl.heff_tensor_indexes(ixf(72))
    for dtype in (np.float64, np.complex128):
        t = {n: tk.BlockSparseTensor(ix, dtype).random(div, rng) for n, ix in ti.items()}
        want = dict(t)
        for l, r, ax, o in wl.HEFF_STEPS:
            want[o] = onp.contract_np(want[l], want[r], ax)
        for flags in (0, _lib.PLAN_CPLX_4M, _lib.PLAN_STAGGER_OUTPUT, _lib.PLAN_NO_SKINNY, _lib.PLAN_PERMUTE_ALL):
            cur = dict(t)
            for l, r, ax, o in wl.HEFF_STEPS:
                m = tk.Match(cur[l], cur[r], ax)
                p = tk.ContractionPlan(ctx, m, dtype, _lib.PLAN_DETERMINISTIC | flags)
                c = m.result_shell(dtype)
                p.execute_host(cur[l].data, cur[r].data, c.data)
                p.close(); m.close()
                cur[o] = c
            assert rel(cur["out"].data, want["out"].data) <= 1e-12, (name, dtype, flags)
        chain = ContractionChain(ctx, t, wl.HEFF_STEPS, dtype)
        chain.make_host_pipe("psi")
        out = np.empty_like(want["out"].data)
        chain.apply_host_pipelined(t["psi"].data, out)
        assert rel(out, want["out"].data) <= 1e-12
        chain.close()
        # accumulate: expansion + scaling of untouched blocks
        t1 = want["t1"]
        half = tk.BlockSparseTensor(t1.indexes, dtype)
        half.set_blocks(t1.blk_coors[::2]); half.data[...] = 1.0
        got = tk.contract_tail_head_contiguous_accumulate(t["lenv"], t["psi"], 0, 0, 1, 0.5, 2.0, half, ctx)
        ref = onp.contract_accumulate_np(t["lenv"], t["psi"], 0, 0, 1, 0.5, 2.0, half)
        assert got.same_structure(ref) and rel(got.data, ref.data) <= 1e-12
        # whole-tensor transposes (every permute mode) and, for bosonic tensors, the axis kernel
        for perm in ([3, 1, 2, 0], [1, 0, 3, 2], [0, 2, 1, 3]):
            a = tk.transpose(t["psi"], perm, ctx)
            b = onp.transpose_np(t["psi"], perm)
            assert np.array_equal(a.data, b.data)
        if name == "u1":
            ix = ixf(72)
            o1 = tk.BlockSparseTensor([ix["ph_in"], ix["ph_out"]], dtype).random((0,), rng)
            o2 = tk.BlockSparseTensor([ix["ph_in"], ix["ph_out"]], dtype).random((2,), rng)
            a = tk.apply_two_rank2_to_axes_preserve_order(t["psi"], o1, 1, o2, 2, ctx)
            b = onp.apply_rank2_axes_np(t["psi"], [(o1, 1), (o2, 2)])
            assert a.same_structure(b) and rel(a.data, b.data) <= 1e-12
# split-K with several pairs per block
for dtype in (np.float64, np.complex128):
    cplx = dtype == np.complex128
    m, n, ks = 40, 100, [900, 500]
    a_shape, b_shape, a_off, b_off, tasks, ad, bd = [], [], [], [], [], [], []
    ao = bo = 0
    acc = np.zeros((m, n), dtype)
    for pi, k in enumerate(ks):
        a = rng.standard_normal((m, k)) + (1j * rng.standard_normal((m, k)) if cplx else 0)
        b = rng.standard_normal((k, n)) + (1j * rng.standard_normal((k, n)) if cplx else 0)
        acc += a @ b
        tasks.append(dict(a_ord=pi, b_ord=pi, c_ord=0, a_off=ao, b_off=bo, c_off=0, m=m, k=k, n=n, sign=1, first=int(pi == 0)))
        a_shape += [m, k]; b_shape += [k, n]; a_off.append(ao); b_off.append(bo)
        ad.append(a.ravel()); bd.append(b.ravel()); ao += a.size; bo += b.size
    plan = tk.RawPlan(ctx, dtype, 2, [0, 1], a_shape, a_off, 2, [0, 1], b_shape, b_off, tasks, m * n, _lib.PLAN_DETERMINISTIC | _lib.PLAN_NO_SKINNY)
    c = np.zeros(m * n, dtype)
    plan.execute_host(np.concatenate(ad).astype(dtype), np.concatenate(bd).astype(dtype), c)
    assert plan.stats().ntile_dmma > 2 and rel(c, acc.ravel()) <= 1e-12
    plan.close()
ctx.sync()
print("sanitizer cases ok")
