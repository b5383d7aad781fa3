#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke exit $rc" >> gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then export R3G_GEMM_2CTA=0; echo "2cta disabled" >> gpurun_out/summary.txt; fi
timeout 600 python -m pytest tests/test_gpu_linear_attention.py tests/test_gpu_models.py tests/test_gpu_vggt.py -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log
R3G_MB_ONLY=linear R3G_MB_OUT=mb_lin_2cta.json timeout 200 python tools/microbench.py > gpurun_out/mb_lin_2cta.log 2>&1; tail -9 gpurun_out/mb_lin_2cta.log | cut -c1-150
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_v3.json | cut -c1-2500; tail -3 gpurun_out/bench_v3.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"linear_kernel|attention_kernel" -o gpurun_out/prof_r1c -f python tools/prof_target.py > gpurun_out/ncu_r1c.log 2>&1; tail -2 gpurun_out/ncu_r1c.log
cat gpurun_out/summary.txt
