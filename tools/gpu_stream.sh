#!/bin/bash
export SVBRDF_B200_QUIET=1
export SVBRDF_B200_LIB=svbrdf_diff_renderer_b200/csrc/libsvbrdf_b200_stream.so
for s in 3 4 6 8 12; do
  echo -n "stream slots=$s per-step:  "; SVBRDF_B200_SLOTS=$s timeout 100 python tools/kernel_bench.py --variants "tma1" --steps 40 2>&1 | grep "^tma1" | cut -c1-90
done
echo -n "stream fused 40 epochs:  "; timeout 100 python tools/kernel_bench.py --variants "tma1" --steps 40 --fused-epochs 2>&1 | grep "^tma1" | cut -c1-90
echo -n "stream u8 targets:       "; timeout 100 python tools/kernel_bench.py --variants "tma1" --steps 40 --u8 2>&1 | grep "^tma1" | cut -c1-90
unset SVBRDF_B200_LIB
echo -n "real   u8 targets:       "; timeout 100 python tools/kernel_bench.py --variants "tma1" --steps 40 --u8 2>&1 | grep "^tma1" | cut -c1-90
echo -n "real   u8 fused epochs:  "; timeout 100 python tools/kernel_bench.py --variants "tma1" --steps 40 --u8 --fused-epochs 2>&1 | grep "^tma1" | cut -c1-90
