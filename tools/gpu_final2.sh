#!/bin/bash
# Final single-GPU visit of session 2: smoke, full GPU test-suite, both bench arms, kernel table, launch list.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-v10}
export SVBRDF_B200_QUIET=1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_$TAG.txt
echo "== bench" ; timeout 600 python bench.py --steps 40 --warmup 5 2>$OUT/bench.err > $OUT/bench_$TAG.json; wc -l $OUT/bench_$TAG.json; cut -c1-300 $OUT/bench_$TAG.json
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>$OUT/bench.err > $OUT/bench_ref_$TAG.json; cut -c1-200 $OUT/bench_ref_$TAG.json
echo "== kernel table"; timeout 600 python tools/kernel_table.py 2>&1 | tee $OUT/kernel_table_$TAG.md | tail -18
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-view-sharded > $OUT/ncu_launch_bench.log 2>&1
grep -c _kernel $OUT/launches_$TAG.csv
echo "== done"
