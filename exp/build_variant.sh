#!/bin/bash
# usage: exp/build_variant.sh NAME "-DFOO=1 -DBAR=2"  -> exp/variants/libqlb200_NAME.so (same sources, extra defines)
set -e
name=$1; defs=$2
src=tensortoolkit_b200/csrc
out=exp/variants; mkdir -p $out/$name
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Iinclude -I$src --expt-relaxed-constexpr $defs"
for f in matcher accum axis shard plan; do g++ -O2 -std=c++17 -fPIC -Iinclude -I$src -I/usr/local/cuda/include $defs -c $src/$f.cc -o $out/$name/$f.o & done
for f in permute gemm gemm_ws gemm_ws_real axis_kernel comm capi; do $NV -c $src/$f.cu -o $out/$name/$f.o & done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libqlb200_$name.so $out/$name/*.o -cudart static
rm -rf $out/$name
echo built $out/libqlb200_$name.so
