#!/bin/bash
# Session 3, visit B: default library after the exact packed subtraction (full GPU suite), chunk-size variants
# (ch4 = 4 scalar lights per slot, pl6 = 3 light pairs per 6-light slot), packed two-texel kernel at a rendered ground truth.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-s3b}
export SVBRDF_B200_QUIET=1
C=svbrdf_diff_renderer_b200/csrc
echo "== pytest -m gpu (default lib)"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_$TAG.txt
echo "== packed two-texel kernel at a rendered ground truth"
SVBRDF_B200_PACKED=1 timeout 120 python - <<'PY' 2>&1 | tail -3 | tee $OUT/packed_fixed_point_$TAG.txt
import torch as th, svbrdf_diff_renderer_b200 as pkg
from svbrdf_diff_renderer_b200 import synth
dev = th.device("cuda:0"); res, n = 512, 9
cl = [c.to(dev) for c in synth.calibration(n)]
r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, cl, dev)
gt = synth.random_textures(res, 1).to(dev)
with th.no_grad(): target = r.eval(gt)
o = pkg.SvbrdfOptim(dev, r); o.load_targets(target); o.init_from_tex(gt.clone())
print("packed: losses at the ground truth", o.optim(3, 0.01, None, False, progress=False), "max |step|", float((o.textures.detach() - gt).abs().max()))
PY
for lib in ${LIBS:-base ch4 pl6}; do
  if [ "$lib" = base ]; then unset SVBRDF_B200_LIB; else export SVBRDF_B200_LIB=$C/libsvbrdf_b200_$lib.so; fi
  echo "== lib $lib 1024x9 (40 epochs per launch)" | tee -a $OUT/variants_$TAG.txt
  timeout 200 python tools/kernel_bench.py --fused-epochs --steps 40 --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
  echo "== lib $lib 2048x64" | tee -a $OUT/variants_$TAG.txt
  timeout 200 python tools/kernel_bench.py --res 2048 --lights 64 --mats 1 --steps 10 --variants "tma1" 2>&1 | grep -v '^{' | tail -1 | tee -a $OUT/variants_$TAG.txt
done
echo "== done"
