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

from .tensor import IN, OUT, Index, QNSector, U1, fU1U1

HEFF_STEPS = [  # (lhs, rhs, axes, out)
    ("lenv", "psi", ([0], [0]), "t1"),
    ("t1", "mpo1", ([0, 2], [0, 1]), "t2"),
    ("t2", "mpo2", ([4, 1], [0, 1]), "t3"),
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
