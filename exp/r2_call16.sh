#!/bin/bash
# round 2, call 16: where does the real-double GEMM lose its 17 %?  Deliberately wrong variants timed only (no copies issued /
# fragments not loaded / no epilogue stores), and the split-K cut length at D=1024.
mkdir -p gpurun_out
export QLB200_BENCH_NO_VERIFY=1
run_bench() {   # tag, args...
  local tag=$1; shift
  timeout 200 python bench.py "$@" --breakdown --no-cpu-baseline --steps 10 > gpurun_out/r2_c16_$tag.json 2> gpurun_out/r2_c16_$tag.err
  echo "== $tag rc=$?"; grep -E "dmma|permute|skinny" gpurun_out/r2_c16_$tag.err | tail -5
}
for v in NOCOPY NOLDS NOEPI; do
  export QLB200_LIB=$PWD/exp/variants/libqlb200_$v.so
  echo "######## $v"
  run_bench d4096f64_$v --D 4096 --dtype f64 --no-cold --no-fused-mpo --no-sub-records
  run_bench ragged_$v --workload ragged
done
unset QLB200_LIB
unset QLB200_BENCH_NO_VERIFY
for c in 12 8 6 4; do
  echo "######## split chunk $c"
  QLB200_SPLIT_CHUNK=$c run_bench d1024f64_chunk$c --D 1024 --dtype f64 --no-cold --no-fused-mpo --no-sub-records
done
QLB200_SPLIT_CHUNK=8 run_bench d1024c128_chunk8 --D 1024 --no-cold --no-fused-mpo --no-sub-records
run_bench d1024c128_default --D 1024 --no-cold --no-fused-mpo --no-sub-records
