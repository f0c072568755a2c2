#!/bin/bash
# ncu --set full of the mode-B kernels at 1024^2 x 9 + the kernel table.   gpurun -- 'TAG=r02 bash tools/gpu_modeb.sh'
set -u
OUT=gpurun_out; TAG=${TAG:-r02}; mkdir -p $OUT
export SVBRDF_B200_QUIET=1
[ -n "${LIB:-}" ] && export SVBRDF_B200_LIB=svbrdf_diff_renderer_b200/csrc/libsvbrdf_b200_$LIB.so
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:texel_kernel|tile_kernel|norm_l2_kernel' -s 5 -c 5 -f -o $OUT/prof_modeb_$TAG \
  python tools/mode_b_run.py > $OUT/ncu_modeb_$TAG.log 2>&1
ls -la $OUT/prof_modeb_$TAG.ncu-rep
python tools/ncu_summary.py $OUT/prof_modeb_$TAG.ncu-rep $OUT/modeb_summary_$TAG.txt > /dev/null 2>&1
echo "== kernel table"; timeout 300 python tools/kernel_table.py | tee $OUT/kernel_table_$TAG.md
echo "== features bench"; timeout 200 python tools/features_bench.py 2>&1 | tail -6 | tee $OUT/features_bench_$TAG.txt
echo "== done"
