#!/bin/bash
export SVBRDF_B200_QUIET=1
echo -n "per-step, 1 material:        "; timeout 200 python tools/kernel_bench.py --variants "tma1" --steps 40 --mats 1 2>&1 | grep "^tma1"
echo -n "run(40), 1 launch/epoch:     "; SVBRDF_B200_EPOCHS_PER_LAUNCH=1 timeout 200 python tools/kernel_bench.py --variants "tma1" --steps 40 --mats 1 --fused-epochs 2>&1 | grep "^tma1"
echo -n "run(40), 8 epochs/launch:    "; SVBRDF_B200_EPOCHS_PER_LAUNCH=8 timeout 200 python tools/kernel_bench.py --variants "tma1" --steps 40 --mats 1 --fused-epochs 2>&1 | grep "^tma1"
echo -n "run(40), 40 epochs/launch:   "; timeout 200 python tools/kernel_bench.py --variants "tma1" --steps 40 --mats 1 --fused-epochs 2>&1 | grep "^tma1"
echo -n "run(40) x4 mats, 40/launch:  "; timeout 200 python tools/kernel_bench.py --variants "tma1" --steps 40 --mats 4 --fused-epochs 2>&1 | grep "^tma1"
echo -n "per-step, 4 materials:       "; timeout 200 python tools/kernel_bench.py --variants "tma1" --steps 40 --mats 4 2>&1 | grep "^tma1"
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv
