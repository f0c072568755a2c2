#!/bin/bash
# GPU visit for the self-fed ring (16 consumer warps, no producer warp): pending_count semantics, parity tests through the
# variant library, A/B against the default build, optional ncu capture.
#   gpurun --timeout 1200 -- 'TAG=r02g LIBS="base self" bash tools/gpu_selffed.sh'
set -u
OUT=gpurun_out; TAG=${TAG:-r02g}; mkdir -p $OUT
export SVBRDF_B200_QUIET=1
C=svbrdf_diff_renderer_b200/csrc
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== mbarrier pending_count"; timeout 60 tools/microbench/mbar_pending | tee $OUT/mbar_pending_$TAG.txt
if ! grep -q "pending: 4 3 2 1 4 3 2 1" $OUT/mbar_pending_$TAG.txt; then echo "UNEXPECTED pending_count semantics: self-fed runs skipped"; exit 0; fi
TESTLIB=${TESTLIB:-self}
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  echo "== pytest -m gpu through lib $TESTLIB"
  SVBRDF_B200_LIB=$C/libsvbrdf_b200_$TESTLIB.so SVBRDF_PARITY_MARGINS=$OUT/parity_margins_$TAG.jsonl timeout 900 python -m pytest tests -m gpu -q -x --durations=5 ${PYTEST_ARGS:-} 2>&1 | tail -25 | tee $OUT/pytest_gpu_$TAG.txt
fi
: > $OUT/variants_$TAG.txt
IFS=';' read -ra CF <<< "${CFGS:---res 1024 --lights 9 --fused-epochs --steps 40 --mats 2;--res 2048 --lights 64 --mats 1 --steps 6;--res 4096 --lights 64 --mats 1 --steps 4}"
for rep in 1 2; do
for lib in ${LIBS:-base self}; do
  if [ "$lib" = base ]; then unset SVBRDF_B200_LIB; else export SVBRDF_B200_LIB=$C/libsvbrdf_b200_$lib.so; fi
  for cfg in "${CF[@]}"; do
    echo "== lib $lib $cfg" | tee -a $OUT/variants_$TAG.txt
    timeout ${KB_TIMEOUT:-120} python tools/kernel_bench.py $cfg --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  done
done
done
unset SVBRDF_B200_LIB
for cfg in ${NCU:-}; do
  RES=${cfg%x*}; LIGHTS=${cfg#*x}
  echo "== ncu --set full $cfg (lib ${NCU_LIB:-self})"
  export SVBRDF_B200_LIB=$C/libsvbrdf_b200_${NCU_LIB:-self}.so
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 6 -c 1 -f -o $OUT/prof_${TAG}_$cfg \
     python tools/kernel_bench.py --res $RES --lights $LIGHTS --mats 1 --steps 3 --variants "tma1" > $OUT/ncu_${TAG}_$cfg.log 2>&1
  unset SVBRDF_B200_LIB
  ls -la $OUT/prof_${TAG}_$cfg.ncu-rep
  ncu -i $OUT/prof_${TAG}_$cfg.ncu-rep --page details --csv > $OUT/prof_${TAG}_${cfg}_details.csv 2>/dev/null
done
echo "== done"
