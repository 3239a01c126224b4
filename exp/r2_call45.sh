#!/bin/bash
# round 2, call 45: final library build: smoke, symbol / introspection tests, a parity subset
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -2 gpurun_out/r2_pytest_gpu.log
