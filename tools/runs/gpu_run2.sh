#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
rm -f gpurun_out/summary.txt
for t in test_gpu_rowops test_gpu_linear_attention test_gpu_models test_gpu_pipeline; do
  timeout 900 python -m pytest tests/$t.py -q -m gpu --timeout 600 > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
  tail -15 gpurun_out/$t.log
done
cat gpurun_out/summary.txt
