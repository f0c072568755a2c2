#!/bin/bash
# Round-1 session-2 visit A: parity of the 10-MUFU light loop, A/B against the 13-MUFU build, ring-depth sweep, packed shape.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-r2a}
export SVBRDF_B200_QUIET=1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $OUT/pytest_gpu_$TAG.txt
for lib in "" ${LIBS:-}; do
  if [ -n "$lib" ]; then export SVBRDF_B200_LIB=svbrdf_diff_renderer_b200/csrc/libsvbrdf_b200_$lib.so; fi
  echo "== lib '${lib:-default}' 1024x9 (40 epochs per launch)" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --fused-epochs --steps 40 --variants "${VARIANTS:-tma1;tma1s8;tma1s7;tma1s6;tma1p}" 2>&1 | grep -v '^{' | tail -8 | tee -a $OUT/variants_$TAG.txt
  echo "== lib '${lib:-default}' 2048x64" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --steps 10 --variants "${VARIANTS64:-tma1;tma1p}" 2>&1 | grep -v '^{' | tail -8 | tee -a $OUT/variants_$TAG.txt
done
echo "== done"
