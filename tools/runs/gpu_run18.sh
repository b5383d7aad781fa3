#!/bin/bash
# mbarrier try_wait with suspend hint: effect on attention flavours and on the GEMM; full parity suite
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
R3G_ATTN=2 R3G_ATTN_VARIANT=3 timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke exit $rc" >> gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "canary failed"; exit 1; fi
for cfg in "2 3" "2 4" "2 2" "4 3" "4 4" "3 0"; do
  set -- $cfg
  R3G_ATTN=$1 R3G_ATTN_VARIANT=$2 R3G_MB_ONLY=attention R3G_MB_OUT=mb_attn_a$1v$2.json timeout 120 python tools/microbench.py > gpurun_out/mb_attn_a$1v$2.log 2>&1
  echo "attn=$1 variant=$2: $(grep -o "'tflops': [0-9.]*" gpurun_out/mb_attn_a$1v$2.log | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
done
R3G_MB_ONLY=linear R3G_MB_OUT=mb_lin_hint.json timeout 200 python tools/microbench.py > gpurun_out/mb_lin_hint.log 2>&1; tail -9 gpurun_out/mb_lin_hint.log | cut -c1-150
R3G_ATTN=2 R3G_ATTN_VARIANT=3 timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest (attn=2 variant=3) exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log
cat gpurun_out/summary.txt
