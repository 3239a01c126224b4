#!/bin/bash
# round 2, call 38: final N=1 records of the shipped build: default bench, parity-at-size tests with their printed figures, smoke
mkdir -p gpurun_out
( time timeout 900 python bench.py --breakdown ) > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -8 gpurun_out/r2_bench.err
( time timeout 600 python -m pytest tests/test_parity_at_size.py -m gpu -q -s ) > gpurun_out/r2_pytest_at_size.log 2>&1; tail -3 gpurun_out/r2_pytest_at_size.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
