#!/bin/bash
# Round-1 GPU evidence run (under gpurun, one B200): parity tests, smoke, both bench arms, ncu launch list,
# ncu --set full of the dominant kernels.  Outputs land in gpurun_out/ and are summarised into profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -6 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-600 gpurun_out/bench_reference.json
timeout 600 python bench.py --D 1024 --dtype f64 --breakdown --no-cpu-baseline > gpurun_out/bench_d1024_f64.json 2> gpurun_out/bench_d1024_f64.err; tail -5 gpurun_out/bench_d1024_f64.err
timeout 600 python bench.py --plan-flags 33 --breakdown --no-cpu-baseline > gpurun_out/bench_4m.json 2> gpurun_out/bench_4m.err; tail -5 gpurun_out/bench_4m.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:GemmWsCplx -s 4 -c 2 -o gpurun_out/gemm_ws_cplx3m -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_full_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:GemmSkinny --launch-skip 7 -c 2 -o gpurun_out/skinny -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_full_skinny.log 2>&1
ls -la gpurun_out | head -40
