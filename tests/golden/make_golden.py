"""Regenerates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libqlref.so,
built from /root/reference by oracle/Makefile) on seeded inputs: operands as produced by
QLTensor::Random after SetRandomSeed, the reference's RawDataCtrctTask list, and the result of
qlten::Contract.  Run from the repo root:  python -m tests.golden.make_golden
"""
import os

import numpy as np

from oracle import refbridge as ref
from tests import util
from tests.golden import io as gio

HERE = os.path.dirname(os.path.abspath(__file__))


def dump(name, idx_a, idx_b, axes, div_a, div_b, dtype, seed):
    a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, seed)
    c = ref.contract(a, b, axes)
    tu, td = ref.contract_tasks(a, b, axes, sorted_by_c=False)
    gio.save_case(os.path.join(HERE, name + ".npz"), a.to_bst(), b.to_bst(), axes, c.to_bst(), tu, td,
                  dict(name=name, seed=seed, div_a=list(div_a), div_b=list(div_b)))
    print(name, "tasks", len(tu), "C elems", c.raw().size, "norm", float(np.linalg.norm(c.raw())))


def main():
    fixed = {(k, n): rest for k, n, *rest in util.fixed_cases() if rest[0][0].sectors[0].dgnc == 3}
    for (kind, name) in [("U1", "3d_2axes_trans"), ("U1", "2d_trace"), ("fU1", "3d_2axes_trans"), ("fU1", "2d_trace"), ("fU1", "3d_first_axis")]:
        idx_a, idx_b, axes, div_a, div_b = fixed[(kind, name)]
        for dtype in (np.float64, np.complex128):
            dump(f"{kind}_{name}_{np.dtype(dtype).name}", idx_a, idx_b, axes, div_a, div_b, dtype, 20260000)
    rng = np.random.default_rng(4242)
    for kind in ("U1U1", "fU1U1", "fZ2", "Z2"):
        idx_a, idx_b, axes, div_a, div_b = util.random_case(kind, rng, rank_a=3, rank_b=3, nctrct=2)
        dump(f"{kind}_random_r3", idx_a, idx_b, axes, div_a, div_b, np.float64, 20260001)


if __name__ == "__main__":
    main()
