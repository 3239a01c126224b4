"""Golden-vector (de)serialisation: one .npz per contraction case dumped from the reference."""
import json

import numpy as np

import tensortoolkit_b200 as tk
from tensortoolkit_b200.tensor import Index, QNSector

KINDS = {k.name: k for k in (tk.U1, tk.fU1, tk.U1U1, tk.fU1U1, tk.Z2, tk.fZ2)}


def idx_to_json(ix):
    return dict(kind=ix.kind.name, dir=ix.dir, sectors=[[list(s.qn), s.dgnc] for s in ix.sectors])


def idx_from_json(d):
    return Index(KINDS[d["kind"]], [QNSector(tuple(q), g) for q, g in d["sectors"]], d["dir"])


def save_case(path, A, B, axes, C, tasks_u, tasks_d, meta):
    np.savez_compressed(
        path, meta=json.dumps(dict(meta, axes=[list(axes[0]), list(axes[1])],
                                   a_idx=[idx_to_json(i) for i in A.indexes], b_idx=[idx_to_json(i) for i in B.indexes],
                                   c_idx=[idx_to_json(i) for i in C.indexes], dtype=A.dtype.name)),
        a_coors=A.blk_coors, a_data=A.data, b_coors=B.blk_coors, b_data=B.data,
        c_coors=C.blk_coors, c_data=C.data, tasks_u=tasks_u, tasks_d=tasks_d)


def load_case(path):
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    dt = np.dtype(meta["dtype"])

    def mk(idx_key, coors, data):
        t = tk.BlockSparseTensor([idx_from_json(d) for d in meta[idx_key]], dt)
        if t.rank:
            t.set_blocks(coors, data)
        else:
            t.data = np.array(data, dtype=dt)
        return t

    return dict(A=mk("a_idx", z["a_coors"], z["a_data"]), B=mk("b_idx", z["b_coors"], z["b_data"]),
                C=mk("c_idx", z["c_coors"], z["c_data"]), axes=(meta["axes"][0], meta["axes"][1]),
                tasks_u=z["tasks_u"], tasks_d=z["tasks_d"], meta=meta)
