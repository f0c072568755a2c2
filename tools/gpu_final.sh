#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
export SVBRDF_B200_QUIET=1
echo "== kernel table"; timeout 600 python tools/kernel_table.py 2>&1 | tee $OUT/kernel_table.md | tail -20
echo "== ncu 4096x64 fused"; RES=4096 LIGHTS=64 TAG=4096x64 bash tools/gpu_profile_cfg.sh 2>&1 | tail -3
