#!/bin/bash
# round 2, call 43: full GPU suite + smoke + short bench on the final build
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
timeout 300 python bench.py --breakdown --no-cpu-baseline --no-sub-records --no-cold --no-fused-mpo --steps 20 > gpurun_out/r2_c43_head.json 2> gpurun_out/r2_c43_head.err; grep -E "dmma|skinny" gpurun_out/r2_c43_head.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_c43_head.json").read().strip().splitlines()[-1])
print("ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 3))
PY
