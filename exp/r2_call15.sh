#!/bin/bash
# round 2, call 15: consumer rewrite of both warp-specialised GEMM kernels (accumulator sign frame instead of per-fragment
# sign flips; full-tile stages with compile-time operand layouts: fragment loads = base + immediate) against the previous build.
mkdir -p gpurun_out
run_bench() {   # tag, args...
  local tag=$1; shift
  timeout 200 python bench.py "$@" --breakdown --no-cpu-baseline --steps 10 > gpurun_out/r2_c15_$tag.json 2> gpurun_out/r2_c15_$tag.err
  echo "== $tag rc=$?"; grep -E "dmma|permute|skinny" gpurun_out/r2_c15_$tag.err | tail -5
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_c15_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("   value", round(d["value"]), d["unit"], "ms", round(d["ms_per_step"], 4), "pct_peak", d.get("pct_fp64_peak"))
except Exception as e:
    print("   no record:", e)
PY
}
( timeout 120 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "ragged_raw or variants" ) > gpurun_out/r2_c15_pytest_a.log 2>&1
tail -3 gpurun_out/r2_c15_pytest_a.log
if ! grep -q " passed" gpurun_out/r2_c15_pytest_a.log || grep -q "failed\|Timeout\|Terminated" gpurun_out/r2_c15_pytest_a.log; then echo "first parity subset did not pass: stopping"; exit 1; fi
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_accumulate.py tests/test_parity_at_size.py -m gpu -x -q ) > gpurun_out/r2_c15_pytest_b.log 2>&1
tail -3 gpurun_out/r2_c15_pytest_b.log
for v in new base; do
  if [ $v != new ]; then export QLB200_LIB=$PWD/exp/variants/libqlb200_$v.so; else unset QLB200_LIB; fi
  echo "######## $v"
  run_bench head_$v --no-sub-records --no-cold --no-fused-mpo
  run_bench d4096f64_$v --D 4096 --dtype f64 --no-cold --no-fused-mpo --no-sub-records
  run_bench hub8192_$v --workload heff_hubbard --no-cold --no-fused-mpo
  run_bench ragged_$v --workload ragged
  run_bench d1024f64_$v --D 1024 --dtype f64 --no-cold --no-fused-mpo --no-sub-records
done
