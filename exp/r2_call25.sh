#!/bin/bash
# round 2, call 25: 16-byte epilogue stores of the real kernel (new) vs before (prev); row-major-A L2 prefetch in the complex kernel (cpf8)
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_accumulate.py tests/test_parity_at_size.py -m gpu -x -q -k "float64 or f64 or double or ragged or hubbard or split_k or transposed or real" ) > gpurun_out/r2_c25_pytest.log 2>&1
tail -2 gpurun_out/r2_c25_pytest.log
run_bench() {   # tag, args...
  local tag=$1; shift
  timeout 200 python bench.py "$@" --breakdown --no-cpu-baseline --steps 10 --no-sub-records --no-cold --no-fused-mpo > gpurun_out/r2_c25_$tag.json 2> gpurun_out/r2_c25_$tag.err
  echo "== $tag rc=$?"; grep -E "dmma" gpurun_out/r2_c25_$tag.err | tail -2
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_c25_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("   ms", round(d["ms_per_step"], 4))
except Exception as e:
    print("   no record:", e)
PY
}
for v in new prev; do
  if [ $v != new ]; then export QLB200_LIB=$PWD/exp/variants/libqlb200_$v.so; else unset QLB200_LIB; fi
  echo "######## $v"
  run_bench d4096f64_$v --D 4096 --dtype f64
  run_bench hub8192_$v --workload heff_hubbard
  run_bench ragged_$v --workload ragged
done
export QLB200_LIB=$PWD/exp/variants/libqlb200_cpf8.so
echo "######## cpf8 (complex)"
run_bench head_cpf8
unset QLB200_LIB
echo "######## new (complex)"
run_bench head_new
