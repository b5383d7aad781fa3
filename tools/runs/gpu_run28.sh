#!/bin/bash
# GEMM epilogue: bias / gate slices staged in shared memory per warp, residual fetched one chunk ahead
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke exit $rc" >> gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "canary failed"; exit 1; fi
timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/summary.txt; tail -12 gpurun_out/pytest_gpu.log
R3G_MB_ONLY=linear R3G_MB_OUT=mb_lin_epi.json timeout 200 python tools/microbench.py > gpurun_out/mb_lin_epi.log 2>&1; cut -c1-140 gpurun_out/mb_lin_epi.log
timeout 300 python tools/ablate_dit.py 2>&1 | grep "^linear\|^attention\|full_ms" | cut -c1-200
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v12.json 2> gpurun_out/bench_v12.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_v12.json | cut -c1-2600; tail -3 gpurun_out/bench_v12.err
cat gpurun_out/summary.txt
