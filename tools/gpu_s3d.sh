#!/bin/bash
# Session 3, visit D: default library with the slot size picked from the light count (3 or 4 lights per ring slot):
# full GPU suite, A/B against SVBRDF_B200_CHUNK=3, bench, kernel table.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-s3d}
export SVBRDF_B200_QUIET=1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $OUT/pytest_gpu_$TAG.txt
for ch in 0 3; do
  export SVBRDF_B200_CHUNK=$ch
  for cfg in "--res 2048 --lights 64 --mats 1 --steps 10" "--res 4096 --lights 64 --mats 1 --steps 5" "--res 1024 --lights 16 --fused-epochs --steps 40" "--res 1024 --lights 9 --fused-epochs --steps 40" "--res 1024 --lights 25 --fused-epochs --steps 40"; do
    echo "== chunk $ch $cfg" | tee -a $OUT/variants_$TAG.txt
    timeout 200 python tools/kernel_bench.py $cfg --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  done
done
unset SVBRDF_B200_CHUNK
echo "== bench" ; timeout 600 python bench.py 2>$OUT/bench.err > $OUT/bench_$TAG.json; wc -l $OUT/bench_$TAG.json; cut -c1-300 $OUT/bench_$TAG.json; tail -3 $OUT/bench.err
echo "== kernel table"; timeout 600 python tools/kernel_table.py 2>&1 | tee $OUT/kernel_table_$TAG.md | tail -18
echo "== done"
