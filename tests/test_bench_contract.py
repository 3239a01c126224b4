"""The driver's contract for `bench.py --impl reference` (the reference's own CPU path; no GPU involved): one JSON line with the
keys the driver reads, under plain python and under torchrun with two ranks (rank 0 alone prints).  Runs at a tiny bond
dimension; needs oracle/_ref (built by __graft_entry__.build() where /root/reference exists)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libqlref.so")), reason="oracle/_ref not built")

KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e"}


def check_line(out: str, n_gpus: int):
    lines = [ln for ln in out.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out                                   # exactly one JSON line, whatever the number of ranks
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "GFLOP/s" and d["value"] > 0 and d["ms_per_step"] > 0 and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@needs_ref
def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1", "--D", "128", "--no-thread-sweep"],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    check_line(out.stdout, 1)


@needs_ref
def test_reference_arm_under_torchrun_prints_once():
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29541", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                          "--D", "128", "--no-thread-sweep"], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    check_line(out.stdout, 2)
