#!/bin/bash
# round 2, call 14: real-double GEMM with 8 consumer warps of 32x32 + 4 producer warps (whole warpgroups for setmaxnreg;
# variants real8 = 104 / 32 registers, real8b = 96 / 48) against the default 4 + 4 build; bulk-copy permute path.
# Every command runs under a short `timeout`: a kernel that spins must not hold the box.
mkdir -p gpurun_out
run_bench() {   # tag, args...
  local tag=$1; shift
  timeout 200 python bench.py "$@" --breakdown --no-cpu-baseline --steps 5 > gpurun_out/r2_c14_$tag.json 2> gpurun_out/r2_c14_$tag.err
  echo "== $tag rc=$?"; grep -E "dmma|permute|skinny" gpurun_out/r2_c14_$tag.err | tail -5
  python - "$tag" <<'EOF'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_c14_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("   value", round(d["value"]), d["unit"], "ms", round(d["ms_per_step"], 4), "pct_peak", d.get("pct_fp64_peak"))
except Exception as e:
    print("   no record:", e)
EOF
}
for v in default real8 real8b; do
  if [ $v != default ]; then export QLB200_LIB=$PWD/exp/variants/libqlb200_$v.so; else unset QLB200_LIB; fi
  echo "######## $v"
  ( timeout 90 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "ragged_raw" ) > gpurun_out/r2_c14_pytest_${v}_a.log 2>&1
  tail -2 gpurun_out/r2_c14_pytest_${v}_a.log
  if ! grep -q " passed" gpurun_out/r2_c14_pytest_${v}_a.log || grep -q "failed\|Timeout\|Terminated" gpurun_out/r2_c14_pytest_${v}_a.log; then echo "$v: first real-GEMM test did not pass (rc above); skipping this variant"; continue; fi
  ( timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_accumulate.py -m gpu -x -q -k "float64 or split_k or transposed or ragged or f64 or double or hubbard" ) > gpurun_out/r2_c14_pytest_${v}_b.log 2>&1
  tail -2 gpurun_out/r2_c14_pytest_${v}_b.log
  run_bench ragged_$v --workload ragged
  run_bench hub8192_$v --workload heff_hubbard --no-cold --no-fused-mpo
  run_bench d4096f64_$v --D 4096 --dtype f64 --no-cold --no-fused-mpo --no-sub-records
  run_bench d1024f64_$v --D 1024 --dtype f64 --no-cold --no-fused-mpo --no-sub-records
done
unset QLB200_LIB
echo "######## permute pass forced (bulk-copy path where runs are 16-byte aligned)"
( timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "transpose or permute_all" ) > gpurun_out/r2_c14_pytest_perm.log 2>&1
tail -2 gpurun_out/r2_c14_pytest_perm.log
run_bench ragged_noview --workload ragged --plan-flags 257
