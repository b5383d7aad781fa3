#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo canary failed; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 2 -c 5 -f -o gpurun_out/prof_gemm_r1f python tools/prof_gemm.py > gpurun_out/ncu_gemm_r1f.log 2>&1
tail -2 gpurun_out/ncu_gemm_r1f.log; ls -la gpurun_out/prof_gemm_r1f.ncu-rep
