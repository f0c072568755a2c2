#!/bin/bash
# ncu --set full of the fused kernel at an arbitrary configuration (RES, LIGHTS), plus timing of kernel variants.
set -u
OUT=gpurun_out; mkdir -p $OUT
RES=${RES:-2048}; LIGHTS=${LIGHTS:-64}; TAG=${TAG:-cfg}
export SVBRDF_B200_QUIET=1
timeout 300 python tools/kernel_bench.py --res $RES --lights $LIGHTS --mats 1 --steps 10 --variants "tma1" 2>&1 | grep -v '^{' | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 6 -c 1 -f -o $OUT/prof_${TAG} \
   python tools/kernel_bench.py --res $RES --lights $LIGHTS --mats 1 --steps 3 --variants "tma1" > $OUT/ncu_${TAG}.log 2>&1
ls -la $OUT/prof_${TAG}.ncu-rep
