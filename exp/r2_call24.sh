#!/bin/bash
# round 2, call 24: L2 prefetch on all four operand paths of the real kernel (new) vs row-major A only (pf8)
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_accumulate.py tests/test_parity_at_size.py -m gpu -x -q -k "float64 or f64 or double or ragged or hubbard or split_k or transposed or real" ) > gpurun_out/r2_c24_pytest.log 2>&1
tail -2 gpurun_out/r2_c24_pytest.log
run_bench() {   # tag, args...
  local tag=$1; shift
  timeout 200 python bench.py "$@" --breakdown --no-cpu-baseline --steps 10 --no-sub-records --no-cold --no-fused-mpo > gpurun_out/r2_c24_$tag.json 2> gpurun_out/r2_c24_$tag.err
  echo "== $tag rc=$?"; grep -E "dmma" gpurun_out/r2_c24_$tag.err | tail -2
}
for v in new pf8; do
  if [ $v != new ]; then export QLB200_LIB=$PWD/exp/variants/libqlb200_$v.so; else unset QLB200_LIB; fi
  echo "######## $v"
  run_bench d4096f64_$v --D 4096 --dtype f64
  run_bench hub8192_$v --workload heff_hubbard
  run_bench ragged_$v --workload ragged
  run_bench d1024f64_$v --D 1024 --dtype f64
done
