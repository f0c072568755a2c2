#!/bin/bash
# For each alternative build: GPU parity subset + timing at the two reference configurations.
set -u
export SVBRDF_B200_QUIET=1
for lib in ${LIBS}; do
  export SVBRDF_B200_LIB=svbrdf_diff_renderer_b200/csrc/libsvbrdf_b200_$lib.so
  echo "== $lib parity"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
  echo -n "== $lib 1024x9 fused40: "; timeout 100 python tools/kernel_bench.py --variants tma1 --steps 40 --fused-epochs 2>&1 | grep "^tma1" | cut -c1-100
  echo -n "== $lib 2048x64:        "; timeout 100 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --steps 10 --variants tma1 2>&1 | grep "^tma1" | cut -c1-100
done
unset SVBRDF_B200_LIB
echo -n "== default 1024x9 fused40: "; timeout 100 python tools/kernel_bench.py --variants tma1 --steps 40 --fused-epochs 2>&1 | grep "^tma1" | cut -c1-100
echo -n "== default 2048x64:        "; timeout 100 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --steps 10 --variants tma1 2>&1 | grep "^tma1" | cut -c1-100
