#!/bin/bash
# round 2, call 3: full GPU suite after the B-view / alt-stream changes; default bench; ragged with and without views
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_call3.log 2>&1
tail -6 gpurun_out/r2_pytest_call3.log
( time python bench.py --breakdown --no-sub-records ) > gpurun_out/r2_bench_call3.json 2> gpurun_out/r2_bench_call3.err
tail -c 600 gpurun_out/r2_bench_call3.err
python bench.py --workload ragged --breakdown --steps 10 > gpurun_out/r2_bench_ragged_view.json 2> gpurun_out/r2_bench_ragged_view.err
tail -3 gpurun_out/r2_bench_ragged_view.err
python bench.py --workload ragged --breakdown --steps 10 --plan-flags 257 --no-cpu-baseline > gpurun_out/r2_bench_ragged_noview.json 2> gpurun_out/r2_bench_ragged_noview.err
tail -3 gpurun_out/r2_bench_ragged_noview.err
