#!/bin/bash
W=$1
mkdir -p gpurun_out; rm -f gpurun_out/shards_$W.log
for r in $(seq 0 $((W-1))); do
  echo "=== rank $r of $W" >> gpurun_out/shards_$W.log
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown --shard-of $W:$r 2>> gpurun_out/shards_$W.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'flops', d['config']['flops_per_step'])" >> gpurun_out/shards_$W.log
done
grep -E "===|ms_per_step|step" gpurun_out/shards_$W.log
