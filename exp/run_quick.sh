#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown 2>&1 >gpurun_out/bench_quick.json | grep -E "step|rror"
cut -c1-330 gpurun_out/bench_quick.json
