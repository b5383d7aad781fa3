#!/bin/bash
# v4 attention (P in TMEM): canary, parity tests, micro-benchmark next to v3 on the same box
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
R3G_ATTN=4 timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_attn4.log 2>&1
rc=$?; echo "smoke v4 exit $rc" >> gpurun_out/summary.txt; tail -2 gpurun_out/smoke_attn4.log
if [ $rc -ne 0 ]; then echo "canary failed"; exit 1; fi
R3G_ATTN=4 timeout 300 python -m pytest tests/test_gpu_linear_attention.py -q -m gpu --timeout 120 > gpurun_out/pytest_attn4.log 2>&1
echo "pytest attn v4 exit $?" >> gpurun_out/summary.txt; tail -6 gpurun_out/pytest_attn4.log
for v in 4 3; do R3G_ATTN=$v R3G_MB_ONLY=attention R3G_MB_OUT=mb_attn_v$v.json timeout 200 python tools/microbench.py > gpurun_out/mb_attn_v$v.log 2>&1; echo "attn v$v"; tail -3 gpurun_out/mb_attn_v$v.log | cut -c1-140; done
cat gpurun_out/summary.txt
