#!/bin/bash
# round 2, call 19: one rank's share of the 8-GPU headline run timed on one GPU (bench.py --shard-of 8:r): default LPT list vs
# stream-K vs forced split-K cut lengths
mkdir -p gpurun_out
run_bench() {   # tag, args...
  local tag=$1; shift
  timeout 200 python bench.py "$@" --breakdown --no-cpu-baseline --steps 10 --no-sub-records --no-cold --no-fused-mpo > gpurun_out/r2_c19_$tag.json 2> gpurun_out/r2_c19_$tag.err
  echo "== $tag rc=$?"; grep -E "dmma|permute|skinny|local steps" gpurun_out/r2_c19_$tag.err | tail -6
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_c19_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("   ms", round(d["ms_per_step"], 4))
except Exception as e:
    print("   no record:", e)
PY
}
for r in 3 0; do
  run_bench lpt_r$r --shard-of 8:$r
  run_bench streamk_r$r --shard-of 8:$r --plan-flags 129
  QLB200_SPLIT_CHUNK=101 run_bench chunk101_r$r --shard-of 8:$r
  QLB200_SPLIT_CHUNK=67 run_bench chunk67_r$r --shard-of 8:$r
done
