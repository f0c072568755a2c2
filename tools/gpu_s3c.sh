#!/bin/bash
# Session 3, visit C: lights per ring slot (scalar light body): base = 3, ch4/ch5/ch6/ch8.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-s3c}
export SVBRDF_B200_QUIET=1
C=svbrdf_diff_renderer_b200/csrc
for lib in ${LIBS:-base ch4 ch5 ch6 ch8}; do
  if [ "$lib" = base ]; then unset SVBRDF_B200_LIB; else export SVBRDF_B200_LIB=$C/libsvbrdf_b200_$lib.so; fi
  echo "== lib $lib 2048x64" | tee -a $OUT/variants_$TAG.txt
  timeout 200 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --steps 10 --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  echo "== lib $lib 4096x64" | tee -a $OUT/variants_$TAG.txt
  timeout 200 python tools/kernel_bench.py --res 4096 --lights 64 --mats 1 --steps 5 --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  echo "== lib $lib 2048x64 u8" | tee -a $OUT/variants_$TAG.txt
  timeout 200 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --steps 10 --u8 --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  echo "== lib $lib 1024x9 (40 epochs per launch)" | tee -a $OUT/variants_$TAG.txt
  timeout 200 python tools/kernel_bench.py --fused-epochs --steps 40 --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  echo "== lib $lib 1024x16 (40 epochs per launch)" | tee -a $OUT/variants_$TAG.txt
  timeout 200 python tools/kernel_bench.py --lights 16 --fused-epochs --steps 40 --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
done
echo "== done"
