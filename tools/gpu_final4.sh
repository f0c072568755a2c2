#!/bin/bash
# Final single-GPU visit of session 3: smoke, GPU suite, both bench arms, launch list, ncu --set full at 4096x64 and 1024x9.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-v13}
export SVBRDF_B200_QUIET=1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $OUT/pytest_gpu_$TAG.txt
echo "== bench" ; timeout 600 python bench.py 2>$OUT/bench.err > $OUT/bench_$TAG.json; wc -l $OUT/bench_$TAG.json; cut -c1-200 $OUT/bench_$TAG.json; tail -3 $OUT/bench.err
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>$OUT/bench.err > $OUT/bench_ref_$TAG.json; cut -c1-200 $OUT/bench_ref_$TAG.json
echo "== ncu launch list"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-view-sharded > $OUT/ncu_launch_bench.log 2>&1
grep -c _kernel $OUT/launches_$TAG.csv
echo "== ncu full 4096x64"; RES=4096 LIGHTS=64 TAG=${TAG}_4096x64 bash tools/gpu_profile_cfg.sh 2>&1 | tail -2
echo "== ncu full 1024x9"; RES=1024 LIGHTS=9 TAG=${TAG}_1024x9 bash tools/gpu_profile_cfg.sh 2>&1 | tail -2
echo "== done"
