#!/bin/bash
# Session 3, visit G: R/G channel pairing on FP32x2 (SV_PAIR_RG=1, default) vs all-scalar channels (norg) with the slot size
# picked per light count — the ncu source view at 4096^2 x 64 shows the packed instructions under math_pipe_throttle.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-s3g}
export SVBRDF_B200_QUIET=1
C=svbrdf_diff_renderer_b200/csrc
for lib in ${LIBS:-base norg}; do  # LIBS="base pm" for the SV_PAIR_MORE A/B
  if [ "$lib" = base ]; then unset SVBRDF_B200_LIB; else export SVBRDF_B200_LIB=$C/libsvbrdf_b200_$lib.so; fi
  for cfg in "--res 2048 --lights 64 --mats 1 --steps 10" "--res 4096 --lights 64 --mats 1 --steps 5" "--res 1024 --lights 16 --fused-epochs --steps 40" "--res 1024 --lights 9 --fused-epochs --steps 40" "--res 1024 --lights 9 --steps 40"; do
    echo "== lib $lib $cfg" | tee -a $OUT/variants_$TAG.txt
    timeout 200 python tools/kernel_bench.py $cfg --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  done
done
echo "== done"
