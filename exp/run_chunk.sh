#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/chunk.log
for r in 1 3 0; do for c in 0 67 51; do
  echo "=== rank $r chunk $c" >> gpurun_out/chunk.log
  QLB200_SPLIT_CHUNK=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown --shard-of 8:$r 2>&1 >/dev/null | grep -E "step [14]" >> gpurun_out/chunk.log
done; done
cat gpurun_out/chunk.log
