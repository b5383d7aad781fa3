#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
R3G_GEMM_2CTA=0 R3G_MB_ONLY=linear R3G_MB_OUT=mb_lin_1cta.json timeout 200 python tools/microbench.py > gpurun_out/mb_lin_1cta.log 2>&1; tail -9 gpurun_out/mb_lin_1cta.log | cut -c1-160
R3G_GEMM_2CTA=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -o gpurun_out/prof_gemm_2cta -f python tools/prof_gemm.py > gpurun_out/ncu_gemm2.log 2>&1; tail -2 gpurun_out/ncu_gemm2.log
R3G_GEMM_2CTA=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -o gpurun_out/prof_gemm_1cta -f python tools/prof_gemm.py > gpurun_out/ncu_gemm1.log 2>&1; tail -2 gpurun_out/ncu_gemm1.log
ls -la gpurun_out | tail -8
