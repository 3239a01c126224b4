#!/bin/bash
# round 2, call 44: L2 prefetch of k x m stored A operands alone (pfat8) vs the shipped build (row-major A only)
mkdir -p gpurun_out
run_bench() {   # tag, args...
  local tag=$1; shift
  timeout 200 python bench.py "$@" --breakdown --no-cpu-baseline --steps 20 --no-sub-records --no-cold --no-fused-mpo > gpurun_out/r2_c44_$tag.json 2> gpurun_out/r2_c44_$tag.err
  echo "== $tag rc=$?"; grep -E "dmma" gpurun_out/r2_c44_$tag.err | tail -2
}
for v in pfat8 default; do
  if [ $v != default ]; then export QLB200_LIB=$PWD/exp/variants/libqlb200_$v.so; else unset QLB200_LIB; fi
  echo "######## $v"
  run_bench d4096f64_$v --D 4096 --dtype f64
  run_bench hub8192_$v --workload heff_hubbard
done
