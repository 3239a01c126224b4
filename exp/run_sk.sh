#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "stream_k or split_k or deterministic" 2>&1 | tail -3
timeout 100 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --breakdown --shard-of 8:1 --plan-flags 129 2>&1 >/dev/null | grep -E "step [14]"
