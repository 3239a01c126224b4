#!/bin/bash
# round 2, call 34: narrow-pair work items per resident CTA slot (8 = default) on one rank's share of the 8-GPU run and on the whole problems
mkdir -p gpurun_out
run_bench() {   # tag, args...
  local tag=$1; shift
  timeout 200 python bench.py "$@" --breakdown --no-cpu-baseline --steps 10 --no-sub-records --no-cold --no-fused-mpo > gpurun_out/r2_c34_$tag.json 2> gpurun_out/r2_c34_$tag.err
  echo "== $tag rc=$?"; grep -E "skinny" gpurun_out/r2_c34_$tag.err | tail -2
}
for n in 8 4 3 2 1; do
  export QLB200_SKINNY_ITEMS_PER_SLOT=$n
  echo "######## items per slot $n"
  run_bench shard0_$n --shard-of 8:0
  run_bench shard3_$n --shard-of 8:3
  run_bench head_$n
  run_bench d1024_$n --D 1024 --dtype f64
  run_bench hubshard_$n --workload heff_hubbard --shard-of 8:2
done
