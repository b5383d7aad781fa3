#!/bin/bash
# 4-CTA cluster GEMM (two CTA pairs sharing the W tile through TMA multicast)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
R3G_DEBUG_GEMM=1 timeout -s KILL 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke (4cta) exit $rc" >> gpurun_out/summary.txt; tail -3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then
  echo "4-CTA canary failed; falling back"; export R3G_GEMM_4CTA=0
  timeout -s KILL 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1 || { echo "fallback canary failed too"; exit 1; }
fi
timeout -s KILL 300 python -m pytest tests/test_gpu_linear_attention.py -q -m gpu --timeout 60 -k linear > gpurun_out/pytest_lin.log 2>&1
echo "pytest linear exit $?" >> gpurun_out/summary.txt; tail -5 gpurun_out/pytest_lin.log
for m in 1 0; do
  R3G_GEMM_4CTA=$m R3G_MB_ONLY=linear R3G_MB_OUT=mb_lin_4cta$m.json timeout -s KILL 200 python tools/microbench.py > gpurun_out/mb_lin_4cta$m.log 2>&1
  echo "R3G_GEMM_4CTA=$m: $(grep -o "'tflops': [0-9.]*" gpurun_out/mb_lin_4cta$m.log | cut -c12-16 | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
done
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/summary.txt; tail -5 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v13.json 2> gpurun_out/bench_v13.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_v13.json | cut -c1-2400; tail -3 gpurun_out/bench_v13.err
cat gpurun_out/summary.txt
