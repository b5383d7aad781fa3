#!/bin/bash
# First GPU pass: build check, per-file test runs (each under its own timeout so a hang cannot take the box),
# then micro-benchmarks.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for t in test_gpu_mc test_gpu_rowops test_gpu_linear_attention; do
  timeout 600 python -m pytest tests/$t.py -q -m gpu -x --timeout 300 > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
  tail -5 gpurun_out/$t.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1; echo "microbench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -20 gpurun_out/microbench.log
