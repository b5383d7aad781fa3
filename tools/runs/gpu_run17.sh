#!/bin/bash
# attention flavours, second sweep + ncu of the best so far (two threads/row, P in TMEM, f32 exps, poly 1/8)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for cfg in "4 3" "4 7" "2 3" "2 4"; do
  set -- $cfg
  R3G_ATTN=$1 R3G_ATTN_VARIANT=$2 timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_a$1v$2.log 2>&1
  rc=$?; echo "smoke attn=$1 variant=$2 exit $rc" >> gpurun_out/summary.txt
  if [ $rc -ne 0 ]; then tail -3 gpurun_out/smoke_a$1v$2.log; continue; fi
  R3G_ATTN=$1 R3G_ATTN_VARIANT=$2 timeout 200 python -m pytest tests/test_gpu_linear_attention.py -q -m gpu --timeout 60 -k attention > gpurun_out/pytest_a$1v$2.log 2>&1
  echo "pytest attn=$1 variant=$2 exit $?" >> gpurun_out/summary.txt
done
for cfg in "4 2" "4 3" "4 4" "4 7" "4 8" "2 2" "2 3" "2 4" "2 5"; do
  set -- $cfg
  R3G_ATTN=$1 R3G_ATTN_VARIANT=$2 R3G_MB_ONLY=attention R3G_MB_OUT=mb_attn_a$1v$2.json timeout 120 python tools/microbench.py > gpurun_out/mb_attn_a$1v$2.log 2>&1
  echo "attn=$1 variant=$2: $(grep -o "'tflops': [0-9.]*" gpurun_out/mb_attn_a$1v$2.log | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
done
R3G_ATTN=4 R3G_ATTN_VARIANT=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention -s 2 -c 1 -f -o gpurun_out/prof_attn_a4v3 python tools/prof_attn.py > gpurun_out/ncu_attn_a4v3.log 2>&1
R3G_ATTN=2 R3G_ATTN_VARIANT=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention -s 2 -c 1 -f -o gpurun_out/prof_attn_a2v3 python tools/prof_attn.py > gpurun_out/ncu_attn_a2v3.log 2>&1
ls -la gpurun_out/*.ncu-rep
cat gpurun_out/summary.txt
