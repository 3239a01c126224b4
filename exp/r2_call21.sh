#!/bin/bash
# round 2, call 21: narrow-pair kernel with descriptor prefetch + two terms per trip (new) / two terms only (sk_noprefetch) / before (base)
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_accumulate.py -m gpu -x -q -k "not variants" ) > gpurun_out/r2_c21_pytest.log 2>&1
tail -2 gpurun_out/r2_c21_pytest.log
run_bench() {   # tag, args...
  local tag=$1; shift
  timeout 200 python bench.py "$@" --breakdown --no-cpu-baseline --steps 10 --no-sub-records --no-cold --no-fused-mpo > gpurun_out/r2_c21_$tag.json 2> gpurun_out/r2_c21_$tag.err
  echo "== $tag rc=$?"; grep -E "skinny" gpurun_out/r2_c21_$tag.err | tail -2
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_c21_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("   ms", round(d["ms_per_step"], 4))
except Exception as e:
    print("   no record:", e)
PY
}
for v in new sk_noprefetch base; do
  if [ $v != new ]; then export QLB200_LIB=$PWD/exp/variants/libqlb200_$v.so; else unset QLB200_LIB; fi
  echo "######## $v"
  run_bench head_$v
  run_bench shard0_$v --shard-of 8:0
  run_bench d1024f64_$v --D 1024 --dtype f64
  run_bench hub8192_$v --workload heff_hubbard
done
