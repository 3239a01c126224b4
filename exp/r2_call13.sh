#!/bin/bash
# round 2, call 13: real-double GEMM with 8 consumer warps of 32x32 + 2 producer warps (default build) vs 4 + 4 (variant real4).
# Every command runs under `timeout`: a kernel that spins must not hold the box.
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "variants or split_k or transposed or ragged_raw or transpose" ) > gpurun_out/r2_pytest_call13a.log 2>&1
tail -4 gpurun_out/r2_pytest_call13a.log
if ! grep -q " passed" gpurun_out/r2_pytest_call13a.log || grep -q "failed\|Timeout\|Terminated" gpurun_out/r2_pytest_call13a.log; then echo "first parity subset did not pass: stopping"; exit 1; fi
( time timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_parity_at_size.py tests/test_accumulate.py -m gpu -x -q -k "not complex_3m" ) > gpurun_out/r2_pytest_call13.log 2>&1
tail -4 gpurun_out/r2_pytest_call13.log
for v in default real4; do
  if [ $v != default ]; then export QLB200_LIB=$PWD/exp/variants/libqlb200_$v.so; else unset QLB200_LIB; fi
  timeout 200 python bench.py --workload heff_hubbard --breakdown --no-cpu-baseline --no-cold --no-fused-mpo --steps 5 > gpurun_out/r2_bench_c13_hub_$v.json 2> gpurun_out/r2_bench_c13_hub_$v.err
  echo "== $v hubbard 8192"; tail -4 gpurun_out/r2_bench_c13_hub_$v.err | grep dmma
  timeout 200 python bench.py --workload heff_hubbard --D 4096 --breakdown --no-cpu-baseline --no-cold --no-fused-mpo --steps 5 > gpurun_out/r2_bench_c13_hub4096_$v.json 2> gpurun_out/r2_bench_c13_hub4096_$v.err
  echo "== $v hubbard 4096"; tail -4 gpurun_out/r2_bench_c13_hub4096_$v.err | grep dmma
  timeout 200 python bench.py --workload ragged --breakdown --no-cpu-baseline --steps 5 > gpurun_out/r2_bench_c13_ragged_$v.json 2> gpurun_out/r2_bench_c13_ragged_$v.err
  echo "== $v ragged"; tail -3 gpurun_out/r2_bench_c13_ragged_$v.err | grep dmma
  timeout 200 python bench.py --D 1024 --dtype f64 --breakdown --no-cpu-baseline --no-cold --no-fused-mpo --steps 10 > gpurun_out/r2_bench_c13_d1024_$v.json 2> gpurun_out/r2_bench_c13_d1024_$v.err
  echo "== $v d1024"; tail -4 gpurun_out/r2_bench_c13_d1024_$v.err | grep dmma
  timeout 200 python bench.py --D 4096 --dtype f64 --breakdown --no-cpu-baseline --no-cold --no-fused-mpo --steps 10 > gpurun_out/r2_bench_c13_d4096f64_$v.json 2> gpurun_out/r2_bench_c13_d4096f64_$v.err
  echo "== $v d4096 f64"; tail -4 gpurun_out/r2_bench_c13_d4096f64_$v.err | grep dmma
done
unset QLB200_LIB
timeout 200 python bench.py --workload ragged --breakdown --steps 5 --plan-flags 257 --no-cpu-baseline > gpurun_out/r2_bench_c13_ragged_noview.json 2> gpurun_out/r2_bench_c13_ragged_noview.err
echo "== ragged, permute pass forced (bulk-copy path where runs are 16-byte aligned)"; tail -3 gpurun_out/r2_bench_c13_ragged_noview.err
