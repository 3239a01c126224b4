#!/bin/bash
# round 2, call 2: accumulate / host pipeline / split / sharding GPU tests, then the default bench (with sub-records)
mkdir -p gpurun_out
( time python -m pytest tests/test_accumulate.py tests/test_sharding.py tests/test_parity_gpu.py -m gpu -x -q -k "accumulate or sharded or fused or idle or pipeline or split or dropin" ) > gpurun_out/r2_pytest_call2.log 2>&1
tail -15 gpurun_out/r2_pytest_call2.log
( time python bench.py --breakdown ) > gpurun_out/r2_bench_call2.json 2> gpurun_out/r2_bench_call2.err
tail -c 2500 gpurun_out/r2_bench_call2.err
