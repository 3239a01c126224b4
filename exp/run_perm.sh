#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --workload ragged --steps 5 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/bench_ragged2.json 2> gpurun_out/bench_ragged2.err; tail -3 gpurun_out/bench_ragged2.err; cut -c1-160 gpurun_out/bench_ragged2.json
