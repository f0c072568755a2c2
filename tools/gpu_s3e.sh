#!/bin/bash
# Session 3, visit E: more resident warps at fewer registers (19+1 warps x 96 regs, 23+1 x 80) with 3- and 4-light slots.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-s3e}
export SVBRDF_B200_QUIET=1
C=svbrdf_diff_renderer_b200/csrc
for lib in ${LIBS:-base cw19 cw23}; do
  if [ "$lib" = base ]; then unset SVBRDF_B200_LIB; else export SVBRDF_B200_LIB=$C/libsvbrdf_b200_$lib.so; fi
  for ch in 3 4; do
    export SVBRDF_B200_CHUNK=$ch
    for cfg in "--res 2048 --lights 64 --mats 1 --steps 10" "--res 1024 --lights 9 --fused-epochs --steps 40"; do
      echo "== lib $lib chunk $ch $cfg" | tee -a $OUT/variants_$TAG.txt
      timeout 200 python tools/kernel_bench.py $cfg --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
    done
  done
done
echo "== done"
