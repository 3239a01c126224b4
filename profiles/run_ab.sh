#!/bin/bash
# A/B: 3M (default) vs 4M complex GEMM, parity tests first.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -8 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
timeout 600 python bench.py --breakdown --plan-flags 33 --no-cpu-baseline > gpurun_out/bench_4m.json 2> gpurun_out/bench_4m.err; tail -8 gpurun_out/bench_4m.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:GemmWsCplx -s 4 -c 2 -o gpurun_out/gemm_ws_cplx3m -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_gemm.log 2>&1
ls -la gpurun_out
