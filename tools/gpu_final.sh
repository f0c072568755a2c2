#!/bin/bash
# Closing single-GPU visit of a round: smoke, GPU tests with margins, bench (both arms), ncu launch list of the bench
# command, ncu --set full of the fused kernel at the two judged configurations.
#   gpurun --timeout 1800 -- 'TAG=r02_final bash tools/gpu_final.sh'
set -u
OUT=gpurun_out; TAG=${TAG:-final}; mkdir -p $OUT
export SVBRDF_B200_QUIET=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total,pcie.link.gen.current,pcie.link.width.current --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke_$TAG.txt
echo "== pytest -m gpu"
rm -f $OUT/parity_margins_$TAG.jsonl
SVBRDF_PARITY_MARGINS=$OUT/parity_margins_$TAG.jsonl timeout 900 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -12 | tee $OUT/pytest_gpu_$TAG.txt
echo "== bench"; timeout 900 python bench.py 2>$OUT/bench_$TAG.err > $OUT/bench_$TAG.json; cut -c1-400 $OUT/bench_$TAG.json
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>$OUT/bench_$TAG.err > $OUT/bench_reference_arm_$TAG.json; cut -c1-400 $OUT/bench_reference_arm_$TAG.json
echo "== ncu launch list of the bench command (main legs + config3)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-view-sharded --no-config5 --no-material-batch --no-eager --no-descriptor > $OUT/bench_under_ncu_$TAG.log 2>&1
grep -c tile_kernel $OUT/launches_$TAG.csv
for cfg in ${NCU:-1024x9 4096x64}; do
  RES=${cfg%x*}; LIGHTS=${cfg#*x}
  echo "== ncu --set full $cfg"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 6 -c 1 -f -o $OUT/prof_${TAG}_$cfg \
     python tools/kernel_bench.py --res $RES --lights $LIGHTS --mats 1 --steps 3 --variants "tma1" > $OUT/ncu_${TAG}_$cfg.log 2>&1
  python tools/ncu_summary.py $OUT/prof_${TAG}_$cfg.ncu-rep $OUT/tile_kernel_${cfg}_summary_$TAG.txt > /dev/null 2>&1
  rm -f $OUT/prof_${TAG}_$cfg.ncu-rep
done
echo "== done"
