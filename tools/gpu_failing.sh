#!/bin/bash
# Re-run selected GPU tests with full tracebacks, for the default library and the variants named in LIBS.
set -u
OUT=gpurun_out; TAG=${TAG:-fail}; mkdir -p $OUT
export SVBRDF_B200_QUIET=1
C=svbrdf_diff_renderer_b200/csrc
for lib in ${LIBS:-base}; do
  if [ "$lib" = base ]; then unset SVBRDF_B200_LIB; else export SVBRDF_B200_LIB=$C/libsvbrdf_b200_$lib.so; fi
  echo "== lib $lib"
  timeout 900 python -m pytest tests -m gpu -q -k "${K}" --tb=short 2>&1 | grep -v "^  /\|Warning\|^$" | cut -c1-600 > $OUT/pytest_${TAG}_$lib.txt
  tail -5 $OUT/pytest_${TAG}_$lib.txt
done
