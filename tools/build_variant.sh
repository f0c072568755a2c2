#!/bin/bash
# Build an alternative library (same ABI) with extra -D switches, next to the default one:
#   tools/build_variant.sh NAME -DSV_PAIR_LIGHTS=1 -DSV_CHUNK_LIGHTS=4
# -> svbrdf_diff_renderer_b200/csrc/libsvbrdf_b200_NAME.so, used with SVBRDF_B200_LIB=... (tools/kernel_bench.py, tests).
# Prints the ptxas register / spill lines of the fused scalar-shape kernels.
set -eu
NAME=$1; shift
C=svbrdf_diff_renderer_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false --shared -Xcompiler -fPIC -diag-suppress 128 \
  -Xptxas -v "$@" -o $C/libsvbrdf_b200_$NAME.so $C/svbrdf_kernels.cu $C/svbrdf_maps.cu 2> /tmp/ptxas_$NAME.log || { tail -30 /tmp/ptxas_$NAME.log; exit 1; }
grep -A2 "tile_kernelILi[123]ELb0ELi[01]ENS_9TileShapeILi15ELi1" /tmp/ptxas_$NAME.log | grep -v "^--" | grep -v "Compiling" | paste - - | sed -e 's/ptxas info    ://g' | cut -c1-260
