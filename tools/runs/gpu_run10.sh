#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
ok=""
for v in 3 2; do
  R3G_ATTN=$v timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_attn$v.log 2>&1
  rc=$?; echo "smoke attn v$v exit $rc" >> gpurun_out/summary.txt; tail -1 gpurun_out/smoke_attn$v.log
  if [ $rc -eq 0 ]; then ok=$v; break; fi
done
if [ -z "$ok" ]; then echo "canaries failed"; cat gpurun_out/summary.txt; exit 1; fi
export R3G_ATTN=$ok
timeout 600 python -m pytest tests/test_gpu_linear_attention.py tests/test_gpu_models.py tests/test_gpu_pipeline.py tests/test_gpu_vggt.py -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest (attn v$ok) exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log
for v in 3 2; do R3G_ATTN=$v R3G_MB_ONLY=attention R3G_MB_OUT=mb_attn_v$v.json timeout 200 python tools/microbench.py > gpurun_out/mb_attn_v$v.log 2>&1; echo "attn v$v"; tail -3 gpurun_out/mb_attn_v$v.log | cut -c1-140; done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_v6.json | cut -c1-2600; tail -3 gpurun_out/bench_v6.err
cat gpurun_out/summary.txt
