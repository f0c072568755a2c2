#!/bin/bash
# One GPU-box visit (round 2): smoke, parity tests with margins, kernel-variant A/B, optional ncu capture.
#   gpurun --timeout 1200 -- 'TAG=r02a LIBS="base v1" bash tools/gpu_visit.sh'
# Env: TAG, LIBS (variant library suffixes for tools/kernel_bench.py; "base" = the default build), CFGS (kernel_bench argument
# sets, ';'-separated), SKIP_TESTS=1, NCU="RESxLIGHTS ..." (ncu --set full of the fused kernel per entry), BENCH=1
set -u
OUT=gpurun_out; TAG=${TAG:-r02}; mkdir -p $OUT
export SVBRDF_B200_QUIET=1
C=svbrdf_diff_renderer_b200/csrc
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  rm -f $OUT/parity_margins_$TAG.jsonl
  echo "== pytest -m gpu"
  SVBRDF_PARITY_MARGINS=$OUT/parity_margins_$TAG.jsonl timeout 1500 python -m pytest tests -m gpu -q --durations=8 ${PYTEST_ARGS:-} 2>&1 | tail -40 | tee $OUT/pytest_gpu_$TAG.txt
fi
if [ -n "${LIBS:-}" ]; then
  : > $OUT/variants_$TAG.txt
  IFS=';' read -ra CF <<< "${CFGS:---res 1024 --lights 9 --fused-epochs --steps 40 --mats 2;--res 2048 --lights 64 --mats 1 --steps 6;--res 4096 --lights 64 --mats 1 --steps 4}"
  for rep in 1 2; do
  for lib in $LIBS; do
    if [ "$lib" = base ]; then unset SVBRDF_B200_LIB; else export SVBRDF_B200_LIB=$C/libsvbrdf_b200_$lib.so; fi
    for cfg in "${CF[@]}"; do
      echo "== lib $lib $cfg" | tee -a $OUT/variants_$TAG.txt
      timeout ${KB_TIMEOUT:-120} python tools/kernel_bench.py $cfg --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
    done
  done
  done
  unset SVBRDF_B200_LIB
fi
for cfg in ${NCU:-}; do
  RES=${cfg%x*}; LIGHTS=${cfg#*x}
  echo "== ncu --set full $cfg (lib ${NCU_LIB:-base})"
  if [ -n "${NCU_LIB:-}" ] && [ "${NCU_LIB}" != base ]; then export SVBRDF_B200_LIB=$C/libsvbrdf_b200_${NCU_LIB}.so; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 6 -c 1 -f -o $OUT/prof_${TAG}_$cfg \
     python tools/kernel_bench.py --res $RES --lights $LIGHTS --mats 1 --steps 3 --variants "tma1" > $OUT/ncu_${TAG}_$cfg.log 2>&1
  unset SVBRDF_B200_LIB
  ls -la $OUT/prof_${TAG}_$cfg.ncu-rep
done
if [ "${BENCH:-0}" = "1" ]; then
  echo "== bench"; timeout 900 python bench.py ${BENCH_ARGS:-} 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json | cut -c1-3000
  tail -5 $OUT/bench_$TAG.err
fi
echo "== done"
