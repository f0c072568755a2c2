#!/bin/bash
# Visit H: texture-segment buffering (2/3/4) x ring depth for tile_kernel_ts.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-r2h}
export SVBRDF_B200_QUIET=1
timeout 120 python tools/kernel_bench.py --res 512 --steps 6 --mats 2 --fused-epochs --variants "tma1;tma1s4;tma1n" 2>&1 | tail -4 | cut -c1-150
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee $OUT/pytest_gpu_$TAG.txt
for lib in default tb2 tb4; do
  if [ "$lib" = default ]; then unset SVBRDF_B200_LIB; else export SVBRDF_B200_LIB=svbrdf_diff_renderer_b200/csrc/libsvbrdf_b200_$lib.so; fi
  echo "== lib $lib 1024x9 (40 epochs per launch)" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --fused-epochs --steps 40 --variants "tma1;tma1s7;tma1s6;tma1s5;tma1n" 2>&1 | grep -v '^{' | tail -8 | tee -a $OUT/variants_$TAG.txt
  echo "== lib $lib 2048x64" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --steps 10 --variants "tma1;tma1n" 2>&1 | grep -v '^{' | tail -8 | tee -a $OUT/variants_$TAG.txt
done
echo "== done"
