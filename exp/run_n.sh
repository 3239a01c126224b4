#!/bin/bash
# usage: run_n.sh N [tag] [extra bench args]  -- bench at N GPUs (torchrun, as the driver launches it)
N=$1; TAG=${2:-n$N}; shift; shift
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --breakdown "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
grep -E "step |rank |rror|Traceback" gpurun_out/bench_$TAG.err | tail -22
python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print('$TAG', d['n_gpus'], 'ms', d['ms_per_step'], 'GF', d['value'], d['config']['parallelism'][:60], d.get('sharded_vs_unsharded_rel_err'), 'e2e', d['e2e']['value'])"
