#!/bin/bash
# Multi-GPU visit: gpurun --gpus N -- 'N=2 TAG=r02n2 bash tools/gpu_multi.sh'
# Runs the multi-GPU tests (2- and 4-rank NCCL workers) and bench.py at N ranks under torchrun.
set -u
OUT=gpurun_out; N=${N:-2}; TAG=${TAG:-multi}; mkdir -p $OUT
export SVBRDF_B200_QUIET=1
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/gpu_$TAG.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  echo "== pytest tests/test_gpu_multi.py"
  SVBRDF_PARITY_MARGINS=$OUT/parity_margins_$TAG.jsonl timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short 2>&1 | tail -25 | cut -c1-400 | tee $OUT/pytest_multi_$TAG.txt
fi
echo "== bench N=$N"
timeout ${BENCH_TIMEOUT:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N ${BENCH_ARGS:-} \
  2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json | cut -c1-600
tail -8 $OUT/bench_$TAG.err | cut -c1-300
echo "== done"
