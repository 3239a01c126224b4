#!/bin/bash
# round 2, call 18: ncu --set full of GemmWsReal (current build) on the U(1) chain at D=4096 double: step 1 (A transposed) and step 4 (A row-major)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:GemmWsReal -c 2 -o gpurun_out/r2_gemm_ws_real_d4096 -f python bench.py --D 4096 --dtype f64 --steps 1 --warmup 1 --no-cpu-baseline --no-cold --no-fused-mpo --no-sub-records --no-graph > /dev/null 2> gpurun_out/r2_ncu_real_d4096.err
tail -3 gpurun_out/r2_ncu_real_d4096.err
ls -la gpurun_out/*.ncu-rep
