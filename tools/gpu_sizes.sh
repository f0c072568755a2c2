export SVBRDF_B200_QUIET=1
for lib in stream DEFAULT; do
 for res in 512 1024 2048 4096; do
  if [ $lib = DEFAULT ]; then unset SVBRDF_B200_LIB; else export SVBRDF_B200_LIB=svbrdf_diff_renderer_b200/csrc/libsvbrdf_b200_$lib.so; fi
  echo -n "$lib res=$res N=9: "; timeout 200 python tools/kernel_bench.py --res $res --lights 9 --mats $([ $res -ge 4096 ] && echo 1 || echo 4) --steps 20 --variants "tma1" 2>&1 | grep "^tma1"
 done
done
