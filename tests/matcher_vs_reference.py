"""Sector matcher: ours (hash join, qlb200_match_create) vs the reference's DataBlkGenForTenCtrct
(O(N_A*N_B) scan) on the block structure of config 4 (fermionic Hubbard H_eff chain; the structure does not
depend on D because every sector keeps degeneracy >= 1).  Host-only, no GPU.

    python tests/matcher_vs_reference.py [D]      (output kept in profiles/r1_matcher_cpu.txt)

Test infrastructure: it calls the oracle (the reference compiled into oracle/_ref) as the thing to compare against."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensortoolkit_b200 as tk
from tensortoolkit_b200 import workloads as wl
from oracle import refbridge as ref

D = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for name, ix in (("U(1) Heisenberg (config 3)", wl.u1_heisenberg_indexes(D)), ("U(1)xU(1) fermionic Hubbard (config 4)", wl.hubbard_indexes(D))):
    ti = wl.heff_tensor_indexes(ix)
    div = (0,) * ti["psi"][0].kind.nvals
    ref.set_seed(1)
    r = {n: ref.RefTensor.new(idxs, np.float64).random(div) for n, idxs in ti.items()}
    t = {n: x.to_bst() for n, x in r.items()}
    print(name, "D =", D)
    import ctypes as C
    from tensortoolkit_b200._lib import lib
    L = ref.lib()
    i32 = lambda v: (C.c_int32 * len(v))(*v)
    i64 = lambda v: (C.c_int64 * len(v))(*v)
    for lhs, rhs, axes, out in wl.HEFF_STEPS:
        sa, sb = t[lhs].shell(), t[rhs].shell()
        aa, ba, n = i32(axes[0]), i32(axes[1]), len(axes[0])
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < 0.5:
            h = C.c_void_p()
            assert lib.qlb200_match_create(sa.ptr(), sb.ptr(), n, aa, ba, C.byref(h)) == 0
            ntask = int(lib.qlb200_match_ntask(h)); lib.qlb200_match_destroy(h); reps += 1
        ours = (time.perf_counter() - t0) / reps
        aa64, ba64 = i64(axes[0]), i64(axes[1])
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < 0.5:      # one run of DataBlkGenForTenCtrct on a default C (count only, nothing copied out)
            nref = int(L.qlref_contract_tasks(r[lhs].h, r[rhs].h, n, aa64, ba64, 0, 0, None, None)); reps += 1
        theirs = (time.perf_counter() - t0) / reps
        assert nref == ntask
        print(f"  {lhs:5s} x {rhs:5s}: A blocks {t[lhs].nblk:5d}  B blocks {t[rhs].nblk:5d}  candidate pairs {t[lhs].nblk * t[rhs].nblk:8d}  tasks {ntask:6d}"
              f"   ours {ours * 1e3:7.3f} ms   reference {theirs * 1e3:8.3f} ms   x{theirs / ours:5.1f}")
        r[out] = ref.contract(r[lhs], r[rhs], axes)
        t[out] = r[out].to_bst()
