#!/bin/bash
# round 2, call 22: L2 prefetch hint on the real kernel's 8-byte cp.async (LDGSTS.E.LTC128B / LTC256B) vs none
mkdir -p gpurun_out
run_bench() {   # tag, args...
  local tag=$1; shift
  timeout 200 python bench.py "$@" --breakdown --no-cpu-baseline --steps 10 --no-sub-records --no-cold --no-fused-mpo > gpurun_out/r2_c22_$tag.json 2> gpurun_out/r2_c22_$tag.err
  echo "== $tag rc=$?"; grep -E "dmma" gpurun_out/r2_c22_$tag.err | tail -2
}
for v in default l2pf128 l2pf256; do
  if [ $v != default ]; then export QLB200_LIB=$PWD/exp/variants/libqlb200_$v.so; else unset QLB200_LIB; fi
  echo "######## $v"
  run_bench d4096f64_$v --D 4096 --dtype f64
  run_bench hub8192_$v --workload heff_hubbard
  run_bench ragged_$v --workload ragged
done
