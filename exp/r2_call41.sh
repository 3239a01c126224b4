#!/bin/bash
# round 2, call 41: split-K fix-up with more loads in flight (row-group loop unrolled 2x complex / 4x real) vs before (fix1)
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_accumulate.py -m gpu -x -q -k "split_k or variants or stagger or accumulate" ) > gpurun_out/r2_c41_pytest.log 2>&1
tail -2 gpurun_out/r2_c41_pytest.log
run_bench() {   # tag, args...
  local tag=$1; shift
  timeout 200 python bench.py "$@" --breakdown --no-cpu-baseline --steps 20 --no-sub-records --no-cold --no-fused-mpo > gpurun_out/r2_c41_$tag.json 2> gpurun_out/r2_c41_$tag.err
  echo "== $tag rc=$?"; grep -E "dmma" gpurun_out/r2_c41_$tag.err | tail -2
}
for v in new fix1; do
  if [ $v != new ]; then export QLB200_LIB=$PWD/exp/variants/libqlb200_$v.so; else unset QLB200_LIB; fi
  echo "######## $v"
  run_bench d1024f64_$v --D 1024 --dtype f64
  run_bench d1024c128_$v --D 1024
  run_bench shard0stag_$v --shard-of 8:0 --plan-flags 65
done
