#!/bin/bash
# Visit B: parity after the render/L2 gamma unification; repeated A/B of scalar vs packed shapes (noise estimate).
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-r2b}
export SVBRDF_B200_QUIET=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $OUT/pytest_gpu_$TAG.txt
for rep in 1 2 3; do
for lib in default ofs0 p11; do
  if [ "$lib" = default ]; then unset SVBRDF_B200_LIB; else export SVBRDF_B200_LIB=svbrdf_diff_renderer_b200/csrc/libsvbrdf_b200_$lib.so; fi
  echo "== rep $rep lib $lib 1024x9 (40 epochs per launch)" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --fused-epochs --steps 40 --variants "tma1;tma1p" 2>&1 | grep -v '^{' | tail -8 | tee -a $OUT/variants_$TAG.txt
  echo "== rep $rep lib $lib 2048x64" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --steps 10 --variants "tma1;tma1p" 2>&1 | grep -v '^{' | tail -8 | tee -a $OUT/variants_$TAG.txt
done
done
echo "== done"
