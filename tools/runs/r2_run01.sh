#!/bin/bash
# round 2, run 1: validate the 128-bit row kernels, the single shipped attention kernel, device guards; re-measure rowops
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/r2_01_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_01_pytest.log
for what in rowops attention linear; do
  R3G_MB_ONLY=$what R3G_MB_OUT=r2_01_mb_$what.json timeout 300 python tools/microbench.py > gpurun_out/r2_01_mb_$what.log 2>&1
  echo "== microbench $what rc=$?"; tail -30 gpurun_out/r2_01_mb_$what.log | cut -c1-220
done
