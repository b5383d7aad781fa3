#!/bin/bash
# final kernels: parity suite (incl. the kept attention generations) + ncu launch list with a launch cap (ncu exits cleanly)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout -s KILL 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke exit $rc" >> gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "canary failed"; exit 1; fi
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 150 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log
timeout 520 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 1 --warmup 0 --profile-mode --dit-steps 1 > gpurun_out/ncu_launchlist.log 2>&1
echo "ncu launch list exit $?" >> gpurun_out/summary.txt; wc -l gpurun_out/launches_r1f.csv; tail -2 gpurun_out/ncu_launchlist.log | cut -c1-300
cat gpurun_out/summary.txt
