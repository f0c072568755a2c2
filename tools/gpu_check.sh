#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list + full capture of the fused kernel.
# Usage (from the build container): gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
export SVBRDF_B200_QUIET=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== build" ; python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== bench" ; timeout 600 python bench.py --steps 40 --warmup 5 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-1500
tail -5 $OUT/bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launch_bench.log 2>&1
  grep -c texel_kernel $OUT/launches.csv
  echo "== ncu full capture of the fused kernel"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:texel_kernel -s 12 -c 2 -f -o $OUT/prof_l2adam \
      python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_bench.log 2>&1
  ls -la $OUT/*.ncu-rep
fi
echo "== done"
