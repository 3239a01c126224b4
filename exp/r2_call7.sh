#!/bin/bash
# round 2, call 7: skinny v2 (one thread per row): parity subset, headline + Hubbard with 2 and 3 resident CTAs, fused-MPO chain
mkdir -p gpurun_out
( time python -m pytest tests/test_parity_gpu.py tests/test_accumulate.py tests/test_contiguous.py -m gpu -x -q ) > gpurun_out/r2_pytest_call7.log 2>&1
tail -4 gpurun_out/r2_pytest_call7.log
for v in default skinny3; do
  if [ $v != default ]; then export QLB200_LIB=$PWD/exp/variants/libqlb200_$v.so; else unset QLB200_LIB; fi
  python bench.py --breakdown --no-sub-records --no-cpu-baseline --no-cold --steps 10 > gpurun_out/r2_bench_c7_$v.json 2> gpurun_out/r2_bench_c7_$v.err
  echo "== $v headline"; tail -4 gpurun_out/r2_bench_c7_$v.err
  python bench.py --workload heff_hubbard --breakdown --no-cpu-baseline --no-cold --steps 5 > gpurun_out/r2_bench_c7_hub_$v.json 2> gpurun_out/r2_bench_c7_hub_$v.err
  echo "== $v hubbard"; tail -4 gpurun_out/r2_bench_c7_hub_$v.err
  python bench.py --D 1024 --dtype f64 --breakdown --no-cpu-baseline --no-cold --steps 10 > gpurun_out/r2_bench_c7_d1024_$v.json 2> gpurun_out/r2_bench_c7_d1024_$v.err
  echo "== $v d1024"; tail -4 gpurun_out/r2_bench_c7_d1024_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_c7_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],4), 'fused', d.get('fused_mpo') and {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['fused_mpo'].items() if k!='how'})
    except Exception as e: print(f,'ERR',e)
PY
