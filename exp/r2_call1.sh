#!/bin/bash
# round 2, call 1: parity at size + baseline numbers after the round-1 -> round-2 host changes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
nproc >> gpurun_out/r2_gpu.txt
( time python -m pytest tests/test_parity_at_size.py -m gpu -x -q -s ) > gpurun_out/r2_pytest_at_size.log 2>&1
tail -5 gpurun_out/r2_pytest_at_size.log
( time python -m pytest tests -m gpu -x -q --deselect tests/test_parity_at_size.py ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu.log
( time python bench.py --breakdown ) > gpurun_out/r2_bench_call1.json 2> gpurun_out/r2_bench_call1.err
tail -c 1500 gpurun_out/r2_bench_call1.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2_bench_reference_call1.json 2> gpurun_out/r2_bench_reference_call1.err
tail -3 gpurun_out/r2_bench_reference_call1.err
