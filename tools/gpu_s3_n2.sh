#!/bin/bash
# Session 3: 2-GPU visit — multi-GPU tests and bench.py under torchrun with the slot size picked per light count.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-v13}
export SVBRDF_B200_QUIET=1
echo "== pytest multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_multi_$TAG.txt
echo "== bench n2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 2>$OUT/bench_n2.err > $OUT/bench_n2_$TAG.json; wc -l $OUT/bench_n2_$TAG.json; cut -c1-250 $OUT/bench_n2_$TAG.json; tail -3 $OUT/bench_n2.err
echo "== done"
