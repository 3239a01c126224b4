#!/bin/bash
# round 2, call 30: chunking of the N=1 host pipeline (cumulative fractions of the streamed input / of the output)
mkdir -p gpurun_out
run() {   # tag, pipe-in, pipe-out
  timeout 200 python bench.py --no-cpu-baseline --steps 10 --no-sub-records --no-cold --no-fused-mpo ${2:+--pipe-in $2} ${3:+--pipe-out $3} > gpurun_out/r2_c30_$1.json 2> gpurun_out/r2_c30_$1.err
  python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_c30_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "device ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "serial", round(d["e2e"]["serial_ms_per_step"], 3))
except Exception as e:
    print(sys.argv[1], "no record:", e)
PY
}
run default
run in6 0.03125,0.09375,0.25,0.4375,0.6875,1.0
run out6 "" 0.3125,0.5625,0.75,0.875,0.96875,1.0
run in6out6 0.03125,0.09375,0.25,0.4375,0.6875,1.0 0.3125,0.5625,0.75,0.875,0.96875,1.0
run in5out5 0.04,0.16,0.4,0.7,1.0 0.35,0.65,0.85,0.96,1.0
run in8out8 0.02,0.06,0.14,0.26,0.42,0.6,0.8,1.0 0.25,0.45,0.62,0.76,0.87,0.94,0.98,1.0
