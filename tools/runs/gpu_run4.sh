#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_linear_attention.py tests/test_gpu_models.py tests/test_gpu_pipeline.py tests/test_gpu_rowops.py -q -m gpu --timeout 300 > gpurun_out/pytest_v2.log 2>&1
echo "pytest_v2 exit $?" >> gpurun_out/summary.txt; tail -15 gpurun_out/pytest_v2.log
R3G_MB_ONLY=attention R3G_MB_OUT=mb_attn_v2.json timeout 300 python tools/microbench.py > gpurun_out/mb_attn_v2.log 2>&1; tail -3 gpurun_out/mb_attn_v2.log
R3G_ATTN_V1=1 R3G_MB_ONLY=attention R3G_MB_OUT=mb_attn_v1.json timeout 300 python tools/microbench.py > gpurun_out/mb_attn_v1.log 2>&1; tail -3 gpurun_out/mb_attn_v1.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_v2.json; tail -3 gpurun_out/bench_v2.err
cat gpurun_out/summary.txt
