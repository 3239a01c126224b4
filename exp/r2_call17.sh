#!/bin/bash
# round 2, call 17: attribution for the complex kernel (headline): no copies issued / no epilogue stores (timing only)
mkdir -p gpurun_out
export QLB200_BENCH_NO_VERIFY=1
run_bench() {   # tag, args...
  local tag=$1; shift
  timeout 200 python bench.py "$@" --breakdown --no-cpu-baseline --steps 10 > gpurun_out/r2_c17_$tag.json 2> gpurun_out/r2_c17_$tag.err
  echo "== $tag rc=$?"; grep -E "dmma|permute|skinny" gpurun_out/r2_c17_$tag.err | tail -5
}
for v in NOCOPY NOEPI; do
  export QLB200_LIB=$PWD/exp/variants/libqlb200_$v.so
  echo "######## $v"
  run_bench head_$v --no-sub-records --no-cold --no-fused-mpo
done
