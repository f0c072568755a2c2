#!/bin/bash
# Session 3, visit J: uint8 target decode through I2FP.F32.U32 (ALU pipe; SV_U8_I2FP=1) instead of I2F.U16 (conversion/XU pipe).
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-s3j}
export SVBRDF_B200_QUIET=1
C=svbrdf_diff_renderer_b200/csrc
for lib in base i2fp; do
  if [ "$lib" = base ]; then unset SVBRDF_B200_LIB; else export SVBRDF_B200_LIB=$C/libsvbrdf_b200_$lib.so; fi
  for cfg in "--res 1024 --lights 9 --fused-epochs --u8 --steps 40 --mats 2" "--res 2048 --lights 64 --mats 1 --u8 --steps 6"; do
    echo "== lib $lib $cfg" | tee -a $OUT/variants_$TAG.txt
    timeout 100 python tools/kernel_bench.py $cfg --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  done
done
echo "== pytest uint8 (lib i2fp)"; timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "uint8 or u8 or lights_per_ring_slot" 2>&1 | tail -2 | tee $OUT/pytest_gpu_$TAG.txt
echo "== done"
