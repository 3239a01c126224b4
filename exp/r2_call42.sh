#!/bin/bash
# round 2, call 42: epilogue descriptors handed to the consumers through shared memory with a unit's last stage (new) vs two dependent
# global loads at the start of the epilogue (noepi_desc)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_accumulate.py tests/test_parity_at_size.py tests/test_sharding.py -m gpu -x -q ) > gpurun_out/r2_c42_pytest.log 2>&1
tail -2 gpurun_out/r2_c42_pytest.log
run_bench() {   # tag, args...
  local tag=$1; shift
  timeout 200 python bench.py "$@" --breakdown --no-cpu-baseline --steps 20 --no-sub-records --no-cold --no-fused-mpo > gpurun_out/r2_c42_$tag.json 2> gpurun_out/r2_c42_$tag.err
  echo "== $tag rc=$?"; grep -E "dmma" gpurun_out/r2_c42_$tag.err | tail -2
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_c42_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("   ms", round(d["ms_per_step"], 4))
except Exception as e:
    print("   no record:", e)
PY
}
for v in new noepi_desc; do
  if [ $v != new ]; then export QLB200_LIB=$PWD/exp/variants/libqlb200_$v.so; else unset QLB200_LIB; fi
  echo "######## $v"
  run_bench d1024f64_$v --D 1024 --dtype f64
  run_bench d1024c128_$v --D 1024
  run_bench head_$v
  run_bench hub8192_$v --workload heff_hubbard
  run_bench ragged_$v --workload ragged
done
