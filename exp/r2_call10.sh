#!/bin/bash
# round 2, call 10 (2 GPUs): bench over the C-ABI comm plumbing vs torch symmetric memory
mkdir -p gpurun_out
TAG=n2_capi bash exp/r2_multi.sh 2 --no-sub-records 2>&1 | tail -10
TAG=n2_torch bash exp/r2_multi.sh 2 --no-sub-records --plumbing torch 2>&1 | tail -4
TAG=n2_hub bash exp/r2_multi.sh 2 --workload heff_hubbard 2>&1 | tail -4
TAG=n2_ragged bash exp/r2_multi.sh 2 --workload ragged 2>&1 | tail -4
