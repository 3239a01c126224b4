"""Synthetic workloads of BASELINE.json: index sets and contraction
sequences only; values come from the caller's RNG (bench) or from the reference (parity tests).

The two-site effective-Hamiltonian apply follows the reference's own recipe,
tests/test_tensor_manipulation/test_ten_ctrct_1sct.cc:253-257:
    t1  = Contract(lenv, psi , {{0},{0}})
    t2  = Contract(t1  , mpo1, {{0,2},{0,1}})
    t3  = Contract(t2  , mpo2, {{4,1},{0,1}})
    out = Contract(t3  , renv, {{4,1},{1,0}})      # same indexes as psi
"""
import math
from typing import Dict, List

import numpy as np

from .tensor import IN, OUT, Index, QNSector, U1, fU1U1

HEFF_STEPS = [  # (lhs, rhs, axes, out)
    ("lenv", "psi", ([0], [0]), "t1"),
    ("t1", "mpo1", ([0, 2], [0, 1]), "t2"),
    ("t2", "mpo2", ([4, 1], [0, 1]), "t3"),
    ("t3", "renv", ([4, 1], [1, 0]), "out"),
]


# The same apply with the two MPO tensors pre-contracted into one two-site operator (they are static across the applies of a
# Lanczos run):  w12 = Contract(mpo1, mpo2, {{3},{0}}) = [wb IN, ph IN, ph OUT, ph IN, ph OUT, wb OUT], and the two
# memory-bound MPO steps become ONE pass over the rank-5 intermediate:
#     t3 = Contract(t1, w12, {{0,2,3},{0,1,3}})      # (a2, b) ++ (p1', p2', w'') -- exactly t3's index order above
# It is the matrix-free MPO application of SURVEY.md section 8f rank 4 expressed with the contraction machinery: t1 blocks
# are read in place (stored k x m), w12 blocks are <= 3 x 3 coefficient matrices, the narrow-pair kernel does the rest.
HEFF_FUSE_PREP = ("mpo1", "mpo2", ([3], [0]), "w12")
HEFF_STEPS_FUSED = [
    ("lenv", "psi", ([0], [0]), "t1"),
    ("t1", "w12", ([0, 2, 3], [0, 1, 3]), "t3"),
    ("t3", "renv", ([4, 1], [1, 0]), "out"),
]


def gaussian_degeneracies(weights: List[float], D: int, centre: int) -> List[int]:
    """degeneracy_j = max(1, floor(D * g_j / sum g)), remainder added to the centre sector."""
    tot = sum(weights)
    deg = [max(1, int(math.floor(D * w / tot))) for w in weights]
    rem = D - sum(deg)
    if rem > 0:
        deg[centre] += rem
    return deg


def u1_heisenberg_indexes(D: int) -> Dict[str, Index]:
    """Config 2/3: U(1) spin-1/2 Heisenberg, QN = 2*Sz; 17 bond sectors, sigma = 2.5."""
    js = list(range(-8, 9))
    deg = gaussian_degeneracies([math.exp(-j * j / (2 * 2.5 ** 2)) for j in js], D, 8)
    vb_out = Index(U1, [QNSector((2 * j,), d) for j, d in zip(js, deg)], OUT)
    ph_out = Index(U1, [QNSector((1,), 1), QNSector((-1,), 1)], OUT)
    wb_out = Index(U1, [QNSector((0,), 3), QNSector((2,), 1), QNSector((-2,), 1)], OUT)
    return dict(vb_out=vb_out, vb_in=vb_out.inverse(), ph_out=ph_out, ph_in=ph_out.inverse(),
                wb_out=wb_out, wb_in=wb_out.inverse())


def hubbard_indexes(D: int) -> Dict[str, Index]:
    """Config 4: fermionic Hubbard, QN = fU1U1QN(N, 2Sz); bond sectors (n, s), n+s even."""
    qns, w = [], []
    for n in range(-6, 7):
        for s in range(-4, 5):
            if (n + s) % 2 == 0:
                qns.append((n, s))
                w.append(math.exp(-n * n / (2 * 2.0 ** 2) - s * s / (2 * 1.5 ** 2)))
    deg = gaussian_degeneracies(w, D, qns.index((0, 0)))
    vb_out = Index(fU1U1, [QNSector(q, d) for q, d in zip(qns, deg)], OUT)
    ph_out = Index(fU1U1, [QNSector((0, 0), 1), QNSector((1, 1), 1), QNSector((1, -1), 1), QNSector((2, 0), 1)], OUT)
    wb_out = Index(fU1U1, [QNSector((0, 0), 2), QNSector((1, 1), 1), QNSector((-1, -1), 1),
                           QNSector((1, -1), 1), QNSector((-1, 1), 1)], OUT)
    return dict(vb_out=vb_out, vb_in=vb_out.inverse(), ph_out=ph_out, ph_in=ph_out.inverse(),
                wb_out=wb_out, wb_in=wb_out.inverse())


def heff_tensor_indexes(ix: Dict[str, Index]) -> Dict[str, List[Index]]:
    """psi[vb IN, ph OUT, ph OUT, vb OUT], lenv[vb OUT, wb OUT, vb IN],
    W[wb IN, ph IN, ph OUT, wb OUT], renv[vb IN, wb IN, vb OUT]   (SURVEY.md section 8d)."""
    return dict(
        psi=[ix["vb_in"], ix["ph_out"], ix["ph_out"], ix["vb_out"]],
        lenv=[ix["vb_out"], ix["wb_out"], ix["vb_in"]],
        mpo1=[ix["wb_in"], ix["ph_in"], ix["ph_out"], ix["wb_out"]],
        mpo2=[ix["wb_in"], ix["ph_in"], ix["ph_out"], ix["wb_out"]],
        renv=[ix["vb_in"], ix["wb_in"], ix["vb_out"]],
    )


def ragged_tables(nC=2500, pairs=4, lo=8, hi=2048, seed=20260005):
    """BASELINE configs[4]: descriptor table built directly, no QLTensor.  SplitMix64 stream;
    every C block has (m, n) log-uniform in [lo, hi] and `pairs` contributing pairs with k log-uniform in [lo, hi];
    A blocks stored rank-3 (k, m1, m2) with m1 the largest divisor of m <= sqrt(m), permutation {1,2,0};
    B blocks stored (n1, k, n2), permutation {1,0,2}."""
    state = [seed & 0xFFFFFFFFFFFFFFFF]

    def splitmix():
        state[0] = (state[0] + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = state[0]
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def loguniform():
        u = (splitmix() >> 11) * (1.0 / (1 << 53))
        return int(min(hi, max(lo, round(float(np.exp(np.log(lo) + u * (np.log(hi) - np.log(lo))))))))

    def split(x):
        d = int(x ** 0.5)
        while x % d:
            d -= 1
        return d, x // d

    a_shape, b_shape, a_off, b_off = [], [], [], []
    tasks = np.zeros(nC * pairs, dtype=np.dtype([("a_blk_idx", "<u8"), ("b_blk_idx", "<u8"), ("c_blk_idx", "<u8"), ("a_off", "<u8"), ("b_off", "<u8"),
                                                 ("c_off", "<u8"), ("a_ord", "<u4"), ("b_ord", "<u4"), ("c_ord", "<u4"), ("m", "<u4"), ("k", "<u4"),
                                                 ("n", "<u4"), ("sign", "i1"), ("first", "u1"), ("pad_", "u1", (2,))], align=True))
    ao = bo = co = 0
    flops = 0.0
    ti = 0
    for c in range(nC):
        m, n = loguniform(), loguniform()
        m1, m2 = split(m)
        n1, n2 = split(n)
        for p in range(pairs):
            k = loguniform()
            t = tasks[ti]
            t["a_blk_idx"] = t["a_ord"] = ti; t["b_blk_idx"] = t["b_ord"] = ti; t["c_blk_idx"] = t["c_ord"] = c
            t["a_off"], t["b_off"], t["c_off"] = ao, bo, co
            t["m"], t["k"], t["n"], t["sign"], t["first"] = m, k, n, 1, 1 if p == 0 else 0
            a_shape.append((k, m1, m2)); b_shape.append((n1, k, n2)); a_off.append(ao); b_off.append(bo)
            ao += m * k; bo += k * n
            flops += 2.0 * m * k * n
            ti += 1
        co += m * n
    return dict(a_shape=np.array(a_shape, np.uint32), b_shape=np.array(b_shape, np.uint32), a_off=np.array(a_off, np.uint64),
                b_off=np.array(b_off, np.uint64), tasks=tasks, a_elems=ao, b_elems=bo, c_elems=co, flops=flops)
