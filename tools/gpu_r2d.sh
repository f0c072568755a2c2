#!/bin/bash
# Visit D: where does tile_kernel_ts lose time?  no-fence / no-store builds + ncu source-level capture.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-r2d}
export SVBRDF_B200_QUIET=1
for lib in default nofence nostore; do
  if [ "$lib" = default ]; then unset SVBRDF_B200_LIB; else export SVBRDF_B200_LIB=svbrdf_diff_renderer_b200/csrc/libsvbrdf_b200_$lib.so; fi
  echo "== lib $lib 1024x9 (40 epochs per launch)" | tee -a $OUT/variants_$TAG.txt
  timeout 300 python tools/kernel_bench.py --fused-epochs --steps 40 --variants "tma1;tma1n" 2>&1 | grep -v '^{' | tail -8 | tee -a $OUT/variants_$TAG.txt
done
unset SVBRDF_B200_LIB
echo "== ncu ts kernel 1024x9"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 6 -c 1 -f -o $OUT/prof_ts_1024x9 \
   python tools/kernel_bench.py --res 1024 --lights 9 --mats 1 --steps 3 --variants "tma1" > $OUT/ncu_ts.log 2>&1
ls -la $OUT/prof_ts_1024x9.ncu-rep
echo "== done"
