#!/bin/bash
# timing-only experiments: swap in experimental builds of the library (scratch copy on the GPU box)
mkdir -p gpurun_out
cp tensortoolkit_b200/libqlb200.so /tmp/orig.so
for e in "$@"; do
  cp exp/libqlb200_$e.so tensortoolkit_b200/libqlb200.so
  echo "=== $e" | tee -a gpurun_out/exp.log
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown 2>&1 >/dev/null | grep -E "step [23]|error|Error" | tee -a gpurun_out/exp.log
done
cp /tmp/orig.so tensortoolkit_b200/libqlb200.so
