#!/bin/bash
# attention: packed-fp32 (FFMA2/FADD2) scale-subtract and polynomial exps
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke exit $rc" >> gpurun_out/summary.txt
if [ $rc -ne 0 ]; then echo "canary failed"; exit 1; fi
for cfg in "2 3" "2 6" "2 7" "2 8" "2 9"; do
  set -- $cfg
  R3G_ATTN=$1 R3G_ATTN_VARIANT=$2 timeout 90 python -m pytest tests/test_gpu_linear_attention.py -q -m gpu --timeout 60 -k attention > gpurun_out/pytest_a$1v$2.log 2>&1
  echo "pytest attn=$1 variant=$2 exit $?" >> gpurun_out/summary.txt
  R3G_ATTN=$1 R3G_ATTN_VARIANT=$2 R3G_MB_ONLY=attention R3G_MB_OUT=mb_attn_a$1v$2.json timeout 120 python tools/microbench.py > gpurun_out/mb_attn_a$1v$2.log 2>&1
  echo "attn=$1 variant=$2: $(grep -o "'tflops': [0-9.]*" gpurun_out/mb_attn_a$1v$2.log | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
