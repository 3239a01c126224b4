#!/bin/bash
# round 2, call 6: axis-ops GPU tests; cuBLAS DGEMM/ZGEMM tensor-pipe utilisation (what "peak" means in pipe terms);
# ncu --set full of GemmWsReal on the Hubbard structure at D=4096
mkdir -p gpurun_out
( time python -m pytest tests/test_axis_ops.py -m gpu -x -q ) > gpurun_out/r2_pytest_axis.log 2>&1
tail -6 gpurun_out/r2_pytest_axis.log
cat > /tmp/cublas_probe.py <<'PY'
import torch
for dt, n in ((torch.float64, 8192), (torch.complex128, 4096)):
    a = torch.randn(n, n, dtype=dt, device="cuda"); b = torch.randn(n, n, dtype=dt, device="cuda")
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
PY
ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__block_size,launch__grid_size,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2_ncu_cublas_pipe.csv python /tmp/cublas_probe.py > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_ncu_cublas_pipe.csv | cut -d, -f5,12-15 | sort -u | head -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:GemmWsReal -c 2 -o gpurun_out/r2_gemm_ws_real_hubbard -f python bench.py --workload heff_hubbard --D 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-cold --no-graph > /dev/null 2> gpurun_out/r2_ncu_real.err
tail -3 gpurun_out/r2_ncu_real.err
python bench.py --workload heff_hubbard --D 4096 --steps 10 --breakdown --no-cpu-baseline --no-cold > gpurun_out/r2_bench_hubbard4096.json 2> gpurun_out/r2_bench_hubbard4096.err
tail -5 gpurun_out/r2_bench_hubbard4096.err
