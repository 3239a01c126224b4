#!/bin/bash
# round 2, call 8: narrow-pair kernel with descriptor fetch / term table amortised over 1 / 4 / 8 sub-chunks per item
mkdir -p gpurun_out
( time python -m pytest tests/test_parity_gpu.py tests/test_accumulate.py -m gpu -x -q ) > gpurun_out/r2_pytest_call8.log 2>&1
tail -4 gpurun_out/r2_pytest_call8.log
for v in default sub1 sub8; do
  if [ $v != default ]; then export QLB200_LIB=$PWD/exp/variants/libqlb200_$v.so; else unset QLB200_LIB; fi
  python bench.py --breakdown --no-sub-records --no-cpu-baseline --no-cold --steps 10 > gpurun_out/r2_bench_c8_$v.json 2> gpurun_out/r2_bench_c8_$v.err
  echo "== $v headline"; tail -4 gpurun_out/r2_bench_c8_$v.err | grep skinny
  python bench.py --workload heff_hubbard --breakdown --no-cpu-baseline --no-cold --steps 5 > gpurun_out/r2_bench_c8_hub_$v.json 2> gpurun_out/r2_bench_c8_hub_$v.err
  echo "== $v hubbard"; tail -4 gpurun_out/r2_bench_c8_hub_$v.err | grep skinny
  python bench.py --D 1024 --dtype f64 --breakdown --no-cpu-baseline --no-cold --steps 10 > gpurun_out/r2_bench_c8_d1024_$v.json 2> gpurun_out/r2_bench_c8_d1024_$v.err
  echo "== $v d1024"; tail -4 gpurun_out/r2_bench_c8_d1024_$v.err | grep skinny
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_c8_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],4), 'fused', d.get('fused_mpo') and {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['fused_mpo'].items() if k in ('ms_per_step','mpo_step_ms','mpo_step_gbs')})
    except Exception as e: print(f,'ERR',e)
PY
