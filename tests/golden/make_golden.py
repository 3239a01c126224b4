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


def dump_contiguous(name, kind, dtype, seed):
    """ContractContiguousAxes<Tail, Head> of the reference on a seeded cyclic-contiguous case (tests/test_contiguous.py)."""
    from tests.test_contiguous import contiguous_case
    rng = np.random.default_rng(seed)
    while True:
        idx_a, idx_b, (a0, b0, n), div_a, div_b = contiguous_case(kind, rng)
        if len(idx_a) >= 3 and len(idx_b) >= 2 and 1 <= n < len(idx_a):
            a, b = util.make_ref_pair(ref, idx_a, idx_b, dtype, div_a, div_b, seed)
            c = ref.contract_contiguous(a, b, a0, b0, n)
            if c.raw().size:
                break
    ra, rb = len(idx_a), len(idx_b)
    axes = ([(a0 + i) % ra for i in range(n)], [(b0 + i) % rb for i in range(n)])
    gio.save_case(os.path.join(HERE, name + ".npz"), a.to_bst(), b.to_bst(), axes, c.to_bst(), np.zeros((0, 9), np.uint64), np.zeros((0, 2)),
                  dict(name=name, seed=seed, div_a=list(div_a), div_b=list(div_b), contiguous=[a0, b0, n]))
    print(name, "contiguous", (a0, b0, n), "ranks", ra, rb, "C elems", c.raw().size)


def dump_file(name, kind, dtype, seed):
    """A tensor file written by the reference (ofstream << tensor) next to the tensor itself."""
    rng = np.random.default_rng(seed)
    idx_a, _, _, div_a, _ = util.random_case(kind, rng, rank_a=3, rank_b=2, nctrct=1)
    ref.set_seed(seed)
    a = ref.RefTensor.new(idx_a, dtype).random(div_a)
    a.write_file(os.path.join(HERE, name + ".qlten"))
    A = a.to_bst()
    gio.save_case(os.path.join(HERE, name + ".npz"), A, A, ([], []), A, np.zeros((0, 9), np.uint64), np.zeros((0, 2)),
                  dict(name=name, seed=seed, div_a=list(div_a), div_b=list(div_a), qlten_file=name + ".qlten"))
    print(name, "file bytes", os.path.getsize(os.path.join(HERE, name + ".qlten")))


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
    for kind, dtype in (("U1", np.complex128), ("fU1U1", np.float64), ("fZ2", np.complex128), ("fU1", np.float64)):
        dump_contiguous(f"contig_{kind}_{np.dtype(dtype).name}", kind, dtype, 20260010)
    for kind, dtype in (("U1", np.float64), ("fU1U1", np.complex128), ("Z2", np.float64)):
        dump_file(f"file_{kind}_{np.dtype(dtype).name}", kind, dtype, 20260020)


if __name__ == "__main__":
    main()
