#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list + full capture of the fused kernel.
# Usage (from the build container): gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
# Env: SKIP_NCU=1, SKIP_TESTS=1, VARIANTS="ldg;tma2;..." (tools/kernel_bench.py), TAG=name for profile files
set -u
OUT=gpurun_out
TAG=${TAG:-run}
mkdir -p $OUT
export SVBRDF_B200_QUIET=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== build" ; python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
echo "== tma sanity (short timeout: a pipeline deadlock must not eat the budget)"
timeout 120 python tools/kernel_bench.py --res 256 --steps 3 --mats 2 --variants "tma2;tma2s4" > $OUT/sanity.log 2>&1
SANITY=$?
tail -4 $OUT/sanity.log
if [ $SANITY -ne 0 ]; then
  echo "!! tile_kernel sanity failed or hung -> forcing the LDG kernel for the rest of this visit"
  export SVBRDF_B200_FORCE_LDG=1
fi
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee $OUT/pytest_gpu_$TAG.txt
fi
if [ -n "${VARIANTS:-}" ]; then
  echo "== kernel variants 1024x9" ; timeout 300 python tools/kernel_bench.py --variants "$VARIANTS" 2>&1 | tail -12 | tee $OUT/variants_1024x9_$TAG.txt
  echo "== kernel variants 2048x64" ; timeout 300 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --steps 10 --variants "$VARIANTS" 2>&1 | tail -12 | tee $OUT/variants_2048x64_$TAG.txt
fi
echo "== bench" ; timeout 600 python bench.py --steps 40 --warmup 5 2>$OUT/bench.err | tee $OUT/bench_$TAG.json | cut -c1-1800
tail -5 $OUT/bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$TAG.csv \
      python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-view-sharded > $OUT/ncu_launch_bench.log 2>&1
  grep -c _kernel $OUT/launches_$TAG.csv
  echo "== ncu full capture of the fused kernel"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_kernel|texel_kernel<3' -s 8 -c 2 -f -o $OUT/prof_l2adam_$TAG \
      python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-view-sharded > $OUT/ncu_full_bench.log 2>&1
  ls -la $OUT/*.ncu-rep
fi
echo "== done"
