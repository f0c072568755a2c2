#!/bin/bash
# 8-GPU visit: bench at N=8, phase split of the 1-D and the 4x2 / 2x4 decompositions.
set -u
OUT=gpurun_out; TAG=${TAG:-r02n8}; mkdir -p $OUT
export SVBRDF_B200_QUIET=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "== bench N=8"
timeout 600 $TR --master-port 29511 bench.py --gpus 8 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json | cut -c1-300
tail -3 $OUT/bench_$TAG.err | cut -c1-300
for sh in 8 4 2; do
  echo "== phase split, light shards per band = $sh"
  RES=4096 LIGHTS=256 SHARDS=$sh timeout 300 $TR --master-port 2952$sh tools/peer_phase_timing.py 2>/dev/null | grep -v "^\*\|OMP_NUM\|^$" | sort | tee -a $OUT/peer_phase_$TAG.txt
done
if [ "${TESTS:-1}" = "1" ]; then
  echo "== pytest tests/test_gpu_multi.py"; timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short 2>&1 | tail -5 | tee $OUT/pytest_multi_$TAG.txt
fi
echo "== done"
