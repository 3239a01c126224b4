#!/bin/bash
mkdir -p gpurun_out
timeout 280 python bench.py --workload ragged --steps 5 --warmup 3 --breakdown > gpurun_out/bench_ragged.json 2> gpurun_out/bench_ragged.err; tail -4 gpurun_out/bench_ragged.err; cut -c1-200 gpurun_out/bench_ragged.json
timeout 280 python bench.py --workload heff_hubbard --steps 5 --warmup 3 --breakdown > gpurun_out/bench_hubbard.json 2> gpurun_out/bench_hubbard.err; tail -6 gpurun_out/bench_hubbard.err; cut -c1-200 gpurun_out/bench_hubbard.json
