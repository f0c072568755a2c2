#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
export SVBRDF_B200_QUIET=1
echo "== packed: parity tests"
SVBRDF_B200_PACKED=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
for P in 0 1; do
  echo "== PACKED=$P 1024x9"; SVBRDF_B200_PACKED=$P timeout 200 python tools/kernel_bench.py --variants "tma1" 2>&1 | grep -v '^{' | tail -2
  echo "== PACKED=$P 2048x64"; SVBRDF_B200_PACKED=$P timeout 200 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --steps 10 --variants "tma1" 2>&1 | grep -v '^{' | tail -2
done
