#!/bin/bash
# Session 3, visit A: pair-of-lights FP32x2 shading (two lights of one texel per packed instruction).
#   base = default build (3 lights per slot, scalar); pl3 = pairs + 1 scalar light per 3-light slot; pl4 = 2 pairs per 4-light slot.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-s3a}
export SVBRDF_B200_QUIET=1
C=svbrdf_diff_renderer_b200/csrc
echo "== sanity pl4 (short timeout)"; SVBRDF_B200_LIB=$C/libsvbrdf_b200_pl4.so timeout 120 python tools/kernel_bench.py --res 256 --steps 3 --mats 2 --variants "tma1" 2>&1 | tail -2 | cut -c1-200
for lib in ${LIBS:-base pl3 pl4}; do
  if [ "$lib" = base ]; then unset SVBRDF_B200_LIB; else export SVBRDF_B200_LIB=$C/libsvbrdf_b200_$lib.so; fi
  echo "== lib $lib 1024x9 (40 epochs per launch)" | tee -a $OUT/variants_$TAG.txt
  timeout 200 python tools/kernel_bench.py --fused-epochs --steps 40 --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  echo "== lib $lib 1024x9 (single-epoch launches)" | tee -a $OUT/variants_$TAG.txt
  timeout 200 python tools/kernel_bench.py --steps 40 --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  echo "== lib $lib 1024x9 u8 (40 epochs per launch)" | tee -a $OUT/variants_$TAG.txt
  timeout 200 python tools/kernel_bench.py --fused-epochs --u8 --steps 40 --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  echo "== lib $lib 2048x64" | tee -a $OUT/variants_$TAG.txt
  timeout 200 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --steps 10 --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  echo "== lib $lib 512x9 (40 epochs per launch)" | tee -a $OUT/variants_$TAG.txt
  timeout 200 python tools/kernel_bench.py --res 512 --fused-epochs --steps 40 --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
done
for lib in ${TESTLIBS:-pl4}; do
  export SVBRDF_B200_LIB=$C/libsvbrdf_b200_$lib.so
  echo "== pytest -m gpu with lib $lib"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/pytest_gpu_${TAG}_$lib.txt
done
echo "== done"
