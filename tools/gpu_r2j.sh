#!/bin/bash
# Visit J: contiguous per-CTA spans (SVBRDF_B200_SPANS=1, default) vs round-robin tiles (=0).
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-r2j}
export SVBRDF_B200_QUIET=1
echo "== sanity"; timeout 120 python tools/kernel_bench.py --res 256 --steps 3 --mats 2 --fused-epochs --variants "tma1;tma1p" 2>&1 | tail -3 | cut -c1-150
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_$TAG.txt
for sp in 1 0; do
  export SVBRDF_B200_SPANS=$sp
  for res in 256 512 1024 2048; do
    echo "== spans $sp ${res}x9 (40 epochs per launch)" | tee -a $OUT/variants_$TAG.txt
    timeout 300 python tools/kernel_bench.py --res $res --fused-epochs --steps 40 --mats 2 --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  done
  echo "== spans $sp 1024x9 single-epoch launches" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --steps 40 --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  echo "== spans $sp 2048x64" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --steps 10 --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
done
echo "== done"
