#!/bin/bash
# Visit E: warp-uniform producers (elected issue) for both tile kernels: sanity, parity, A/B.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-r2f}
export SVBRDF_B200_QUIET=1
echo "== sanity"
timeout 120 python tools/kernel_bench.py --res 256 --steps 3 --mats 2 --variants "tma1;tma1n;tma1p" 2>&1 | tail -4 | cut -c1-150
timeout 120 python tools/kernel_bench.py --res 512 --steps 6 --mats 2 --fused-epochs --variants "tma1;tma1s4;tma1n" 2>&1 | tail -4 | cut -c1-150
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $OUT/pytest_gpu_$TAG.txt
for rep in 1 2; do
  echo "== rep $rep 1024x9 (40 epochs per launch)" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --fused-epochs --steps 40 --variants "${VARIANTS:-tma1;tma1n;tma1s7;tma1p}" 2>&1 | grep -v '^{' | tail -8 | tee -a $OUT/variants_$TAG.txt
  echo "== rep $rep 1024x9 (single-epoch launches)" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --steps 40 --variants "tma1;tma1n" 2>&1 | grep -v '^{' | tail -8 | tee -a $OUT/variants_$TAG.txt
  echo "== rep $rep 2048x64" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --steps 10 --variants "tma1;tma1n;tma1p" 2>&1 | grep -v '^{' | tail -8 | tee -a $OUT/variants_$TAG.txt
done
echo "== done"
