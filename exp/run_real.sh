#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown 2>&1 >gpurun_out/bench_quick.json | grep -E "step|rror"
python -c "import json;d=json.load(open('gpurun_out/bench_quick.json'));print('c128 D4096 ms', d['ms_per_step'], 'GF', d['value'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown --dtype f64 2>&1 >gpurun_out/bench_f64_4096.json | grep -E "step|rror"
python -c "import json;d=json.load(open('gpurun_out/bench_f64_4096.json'));print('f64 D4096 ms', d['ms_per_step'], 'GF', d['value'], d['fp64_peak_tflops'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:GemmWsReal -s 2 -c 1 -o gpurun_out/gemm_ws_real -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --dtype f64 > gpurun_out/ncu_real.log 2>&1
