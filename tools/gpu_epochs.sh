#!/bin/bash
set -u
export SVBRDF_B200_QUIET=1
echo "== parity tests (multi-epoch launches inside SvbrdfOptim.optim)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
echo "== per-step launches"; timeout 200 python tools/kernel_bench.py --variants "tma1" --steps 40 2>&1 | grep "^tma1"
echo "== fused epochs (one launch per 40 epochs)"; timeout 200 python tools/kernel_bench.py --variants "tma1" --steps 40 --fused-epochs 2>&1 | grep "^tma1"
echo "== fused epochs 512"; timeout 200 python tools/kernel_bench.py --res 512 --variants "tma1" --steps 40 --fused-epochs 2>&1 | grep "^tma1"
echo "== fused epochs 256"; timeout 200 python tools/kernel_bench.py --res 256 --variants "tma1" --steps 40 --fused-epochs 2>&1 | grep "^tma1"
echo "== fused epochs 2048x64"; timeout 200 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --variants "tma1" --steps 10 --fused-epochs 2>&1 | grep "^tma1"
