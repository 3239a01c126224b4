#!/bin/bash
# round 2, call 48: the default bench line of the final library (record for profiles/r2_bench.json)
mkdir -p gpurun_out
( time timeout 600 python bench.py --breakdown ) > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -6 gpurun_out/r2_bench.err
