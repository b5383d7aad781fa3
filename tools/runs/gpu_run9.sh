#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke exit $rc" >> gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "canary failed"; exit 1; fi
timeout 600 python -m pytest tests/test_gpu_linear_attention.py tests/test_gpu_models.py tests/test_gpu_pipeline.py -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log
R3G_MB_ONLY=linear R3G_MB_OUT=mb_v5.json timeout 200 python tools/microbench.py > gpurun_out/mb_v5.log 2>&1; tail -9 gpurun_out/mb_v5.log | cut -c1-150
R3G_DEBUG_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; echo "bench exit $?" >> gpurun_out/summary.txt
grep "r3g mc" gpurun_out/bench_v5.json | tail -4; grep -v "r3g mc" gpurun_out/bench_v5.json | cut -c1-2600; tail -3 gpurun_out/bench_v5.err
timeout 300 python tools/bench_vggt.py > gpurun_out/bench_vggt.json 2> gpurun_out/bench_vggt.err; cat gpurun_out/bench_vggt.json; tail -3 gpurun_out/bench_vggt.err
cat gpurun_out/summary.txt
