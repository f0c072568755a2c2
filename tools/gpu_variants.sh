#!/bin/bash
# Compare alternative builds (SVBRDF_B200_LIB) and launch shapes of the fused kernel.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-var}
export SVBRDF_B200_QUIET=1
for lib in ${LIBS:-cw7 cw15 cw11 cw7nonr}; do
  echo "== lib $lib 1024x9"
  SVBRDF_B200_LIB=svbrdf_diff_renderer_b200/csrc/libsvbrdf_b200_$lib.so timeout 200 python tools/kernel_bench.py --variants "${VARIANTS:-tma1;tma2;tma3}" 2>&1 | grep -v '^{' | tail -6 | tee -a $OUT/variants_${TAG}.txt
  echo "== lib $lib 2048x64"
  SVBRDF_B200_LIB=svbrdf_diff_renderer_b200/csrc/libsvbrdf_b200_$lib.so timeout 200 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --steps 10 --variants "${VARIANTS:-tma1;tma2;tma3}" 2>&1 | grep -v '^{' | tail -6 | tee -a $OUT/variants_${TAG}.txt
done
