#!/bin/bash
# usage: r2_multi.sh N [extra bench args]  -- torchrun launch exactly as the driver does, bounded by timeout
N=$1; shift
mkdir -p gpurun_out
tag=${TAG:-n$N}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --breakdown "$@" > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err
echo "rc=$?"
tail -c 3000 gpurun_out/r2_bench_$tag.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_$tag.json').read().strip().splitlines()[-1])
    print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',d['e2e'] and (round(d['e2e']['value']),round(d['e2e']['ms_per_step'],3)), 'err', d.get('sharded_vs_unsharded_rel_err'), d['e2e'] and d['e2e'].get('rel_err_vs_device_resident'))
    for k,v in (d.get('sub_records') or {}).items():
        print(' sub',k, v and (round(v['value']), round(v['ms_per_step'],3), v.get('sharded_vs_unsharded_rel_err')))
except Exception as e: print('no json', e)
PY
