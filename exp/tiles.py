import sys, os, numpy as np
sys.path.insert(0, '.')
os.environ["QLB200_DEBUG_TILES"] = "1"
import tensortoolkit_b200 as tk
from tensortoolkit_b200 import workloads as wl
from tensortoolkit_b200.sharding import shard_chain
W = int(sys.argv[1]); D = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ti = wl.heff_tensor_indexes(wl.u1_heisenberg_indexes(D))
t = {n: tk.BlockSparseTensor(ix, np.complex128) for n, ix in ti.items()}
for n in t: t[n].random((0,), np.random.default_rng(1)) if D <= 512 else t[n].set_structure((0,)) if hasattr(t[n], "set_structure") else t[n].random((0,), np.random.default_rng(1))
for r in range(W):
    mine, info = shard_chain(t, wl.HEFF_STEPS, "lenv", 2, W, r, np.complex128) if W > 1 else (t, None)
    sh = dict(mine)
    print(f"--- rank {r}: ranges {[x for x in (info.sector_ranges[r] if info else [])]}", file=sys.stderr)
    for lhs, rhs, axes, out in wl.HEFF_STEPS:
        m = tk.Match(sh[lhs], sh[rhs], axes)
        sh[out] = m.result_shell(np.complex128)
        p = tk.ContractionPlan(None, m, np.complex128)
        p.close(); m.close()
