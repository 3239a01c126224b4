#!/bin/bash
# round 2, call 29: real kernel with the cross-stage software pipeline (swpipe: next stage's first fragments loaded under this stage's
# last k4 step after a non-blocking barrier test) vs default
mkdir -p gpurun_out
( QLB200_LIB=$PWD/exp/variants/libqlb200_swpipe.so timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_accumulate.py tests/test_parity_at_size.py -m gpu -x -q -k "float64 or f64 or double or ragged or hubbard or split_k or transposed or real" ) > gpurun_out/r2_c29_pytest.log 2>&1
tail -2 gpurun_out/r2_c29_pytest.log
run_bench() {   # tag, args...
  local tag=$1; shift
  timeout 200 python bench.py "$@" --breakdown --no-cpu-baseline --steps 10 --no-sub-records --no-cold --no-fused-mpo > gpurun_out/r2_c29_$tag.json 2> gpurun_out/r2_c29_$tag.err
  echo "== $tag rc=$?"; grep -E "dmma" gpurun_out/r2_c29_$tag.err | tail -2
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_c29_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("   ms", round(d["ms_per_step"], 4))
except Exception as e:
    print("   no record:", e)
PY
}
for v in swpipe default; do
  if [ $v != default ]; then export QLB200_LIB=$PWD/exp/variants/libqlb200_$v.so; else unset QLB200_LIB; fi
  echo "######## $v"
  run_bench d4096f64_$v --D 4096 --dtype f64
  run_bench hub8192_$v --workload heff_hubbard
  run_bench ragged_$v --workload ragged
  run_bench d1024f64_$v --D 1024 --dtype f64
done
