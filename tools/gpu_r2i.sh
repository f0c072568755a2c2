#!/bin/bash
# Visit I: per-warp epoch finalisation (no CTA barrier) vs the previous build, same box.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-r2i}
export SVBRDF_B200_QUIET=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_$TAG.txt
for rep in 1 2; do
for lib in default old; do
  if [ "$lib" = default ]; then unset SVBRDF_B200_LIB; else export SVBRDF_B200_LIB=svbrdf_diff_renderer_b200/csrc/libsvbrdf_b200_$lib.so; fi
  echo "== rep $rep lib $lib 1024x9 (40 epochs per launch)" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --fused-epochs --steps 40 --variants "tma1" 2>&1 | grep -v '^{' | tail -2 | tee -a $OUT/variants_$TAG.txt
  echo "== rep $rep lib $lib 1024x9 (single-epoch launches)" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --steps 40 --variants "tma1" 2>&1 | grep -v '^{' | tail -2 | tee -a $OUT/variants_$TAG.txt
  echo "== rep $rep lib $lib 512x9 (40 epochs per launch)" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --res 512 --fused-epochs --steps 40 --variants "tma1" 2>&1 | grep -v '^{' | tail -2 | tee -a $OUT/variants_$TAG.txt
  echo "== rep $rep lib $lib 2048x64" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --steps 10 --variants "tma1" 2>&1 | grep -v '^{' | tail -2 | tee -a $OUT/variants_$TAG.txt
done
done
echo "== done"
