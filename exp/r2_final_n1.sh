#!/bin/bash
# Round-2 evidence run on one B200 (under gpurun): full GPU test suite, smoke, both bench arms, ncu launch list, ncu --set full of
# the dominant kernels, compute-sanitizer on small cases of every kernel family.  Outputs land in gpurun_out/ (r2_*) and are
# summarised into profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt
nproc >> gpurun_out/r2_gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/r2_gpu.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
( time timeout 900 python bench.py --breakdown ) > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -12 gpurun_out/r2_bench.err; head -c 400 gpurun_out/r2_bench.json; echo
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; head -c 700 gpurun_out/r2_bench_reference.json; echo; tail -3 gpurun_out/r2_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub-records --no-cold --no-fused-mpo --no-graph > gpurun_out/r2_ncu_launch.log 2>&1
grep -c "Gemm\|Permute" gpurun_out/r2_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:GemmWsCplx -s 2 -c 2 -o gpurun_out/r2_gemm_ws_cplx3m -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub-records --no-cold --no-fused-mpo --no-graph > gpurun_out/r2_ncu_full_gemm.log 2>&1; tail -2 gpurun_out/r2_ncu_full_gemm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:GemmSkinny -s 2 -c 2 -o gpurun_out/r2_skinny -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub-records --no-cold --no-fused-mpo --no-graph > gpurun_out/r2_ncu_full_skinny.log 2>&1; tail -2 gpurun_out/r2_ncu_full_skinny.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"PermuteKernel|GemmWsReal" -c 2 -o gpurun_out/r2_ragged_permute_gemm -f python bench.py --workload ragged --plan-flags 257 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_full_ragged.log 2>&1; tail -2 gpurun_out/r2_ncu_full_ragged.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:GemmWsReal -c 2 -o gpurun_out/r2_gemm_ws_real_d4096 -f python bench.py --D 4096 --dtype f64 --steps 1 --warmup 1 --no-cpu-baseline --no-cold --no-fused-mpo --no-sub-records --no-graph > gpurun_out/r2_ncu_full_real.log 2>&1; tail -2 gpurun_out/r2_ncu_full_real.log
for tool in memcheck racecheck synccheck; do
  ( time timeout 900 compute-sanitizer --tool $tool python exp/sanitizer_cases.py ) > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer cases ok|Error|hazard" gpurun_out/r2_sanitizer_$tool.log | head -5
done
ls -la gpurun_out | grep r2_ | tail -30
