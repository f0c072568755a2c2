#!/bin/bash
C=svbrdf_diff_renderer_b200/csrc
mkdir -p gpurun_out
python tools/dev/debug_v2.py 2>&1 | tee gpurun_out/dbg_base.txt
SVBRDF_B200_LIB=$C/libsvbrdf_b200_v1.so python tools/dev/debug_v2.py 2>&1 | tee gpurun_out/dbg_v1.txt
