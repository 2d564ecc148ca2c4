#!/bin/bash
# build_variant.sh NAME "-DFLAG=1 ..." -- compiles treelet-prefetching-for-rt_b200/libvsrt_NAME.so with extra nvcc flags (here, no GPU needed);
# the file travels to the GPU box with the snapshot and is selected with VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt_NAME.so
cd "$(dirname "$0")/.." || exit 1
P=treelet-prefetching-for-rt_b200
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC -shared $2 \
  -Iinclude -I$P/csrc $P/csrc/*.cu -o $P/libvsrt_$1.so && echo "built $P/libvsrt_$1.so [$2]"
