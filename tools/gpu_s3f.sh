#!/bin/bash
# Session 3, visit F: probe the next slot's `full` barrier ahead of time (SV_PROBE_AHEAD=1, default build) vs not (noprobe).
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-s3f}
export SVBRDF_B200_QUIET=1
C=svbrdf_diff_renderer_b200/csrc
echo "== sanity (short timeout)"; timeout 120 python tools/kernel_bench.py --res 256 --steps 3 --mats 2 --variants "tma1" 2>&1 | tail -1 | cut -c1-200
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $OUT/pytest_gpu_$TAG.txt
for rep in 1 2; do
for lib in ${LIBS:-base noprobe}; do
  if [ "$lib" = base ]; then unset SVBRDF_B200_LIB; else export SVBRDF_B200_LIB=$C/libsvbrdf_b200_$lib.so; fi
  for cfg in "--res 2048 --lights 64 --mats 1 --steps 10" "--res 1024 --lights 9 --fused-epochs --steps 40" "--res 1024 --lights 9 --steps 40" "--res 1024 --lights 16 --fused-epochs --steps 40" "--res 512 --lights 9 --fused-epochs --steps 40"; do
    echo "== rep $rep lib $lib $cfg" | tee -a $OUT/variants_$TAG.txt
    timeout 200 python tools/kernel_bench.py $cfg --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  done
done
done
echo "== done"
