#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke exit $rc" >> gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "canary failed"; exit 1; fi
timeout 600 python -m pytest tests/test_gpu_linear_attention.py tests/test_gpu_models.py tests/test_gpu_pipeline.py tests/test_gpu_mc.py -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log
R3G_MB_ONLY=linear,attention R3G_MB_OUT=mb_v4.json timeout 200 python tools/microbench.py > gpurun_out/mb_v4.log 2>&1; tail -12 gpurun_out/mb_v4.log | cut -c1-150
timeout 120 python tools/mc_probe.py > gpurun_out/mc_probe.log 2>&1; cat gpurun_out/mc_probe.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_v4.json | cut -c1-2600; tail -3 gpurun_out/bench_v4.err
cat gpurun_out/summary.txt
