#!/bin/bash
# round 2, call 40: permute pass with 8-byte-aligned device buffers (bulk path must step aside), permute / transpose tests
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "8_byte or transpose or permute_all or ragged_raw" ) > gpurun_out/r2_c40_pytest.log 2>&1
tail -5 gpurun_out/r2_c40_pytest.log
timeout 120 compute-sanitizer --tool memcheck python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "8_byte" > gpurun_out/r2_c40_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_c40_memcheck.log | tail -3
