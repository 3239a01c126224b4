#!/bin/bash
# round 2, call 9 (2 GPUs): the C-ABI-only comm test, then the torchrun bench with the partitioner feedback
mkdir -p gpurun_out
( timeout 200 tests/cpp/comm_ranks 1; timeout 200 tests/cpp/comm_ranks 2 ) > gpurun_out/r2_comm_ranks.log 2>&1
cat gpurun_out/r2_comm_ranks.log | tail -8
TAG=n2_fb bash exp/r2_multi.sh 2 --no-sub-records 2>&1 | tail -12
