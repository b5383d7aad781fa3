#!/bin/bash
# Guarded run: 90 s canaries decide which of the new kernels are usable before anything long runs.
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
ok=""
for cfg in "R3G_GEMM_2CTA=1 R3G_ATTN_V1=0" "R3G_GEMM_2CTA=0 R3G_ATTN_V1=0" "R3G_GEMM_2CTA=1 R3G_ATTN_V1=1" "R3G_GEMM_2CTA=0 R3G_ATTN_V1=1"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1
  rc=$?; echo "smoke [$cfg] exit $rc" >> gpurun_out/summary.txt; tail -2 gpurun_out/smoke_$tag.log
  if [ $rc -eq 0 ]; then ok="$cfg"; break; fi
done
if [ -z "$ok" ]; then echo "all canaries failed"; cat gpurun_out/summary.txt; exit 1; fi
export $ok
echo "using: $ok" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/summary.txt; tail -15 gpurun_out/pytest_gpu.log
R3G_MB_OUT=mb_v2.json timeout 300 python tools/microbench.py > gpurun_out/mb_v2.log 2>&1; tail -16 gpurun_out/mb_v2.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_v2.json; tail -3 gpurun_out/bench_v2.err
cat gpurun_out/summary.txt
