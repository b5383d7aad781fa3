#!/bin/bash
# Guarded run: a 90 s canary of the new kernels first; if it hangs or fails, stop (do not burn the budget).
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke exit $rc" >> gpurun_out/summary.txt; tail -3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then
  echo "canary failed: trying attention v1"; export R3G_ATTN_V1=1
  timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v1.log 2>&1 || { echo "v1 canary failed too"; cat gpurun_out/summary.txt; exit 1; }
fi
timeout 600 python -m pytest tests -q -m gpu --timeout 120 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/summary.txt; tail -12 gpurun_out/pytest_gpu.log
R3G_MB_OUT=mb_v2.json timeout 300 python tools/microbench.py > gpurun_out/mb_v2.log 2>&1; tail -16 gpurun_out/mb_v2.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_v2.json; tail -3 gpurun_out/bench_v2.err
cat gpurun_out/summary.txt
