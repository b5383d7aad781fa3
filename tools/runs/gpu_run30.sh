#!/bin/bash
# r1f evidence pass: full parity suite, bench line (with cpu_baseline), reference arm, short ncu launch list
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout -s KILL 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke exit $rc" >> gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "canary failed"; exit 1; fi
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_final.json | cut -c1-3200; tail -3 gpurun_out/bench_final.err
timeout -s KILL 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "reference arm exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_reference_arm.json | cut -c1-900
timeout -s KILL 560 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 1 --warmup 0 --profile-mode --dit-steps 2 > gpurun_out/ncu_launchlist.log 2>&1
echo "ncu launch list exit $?" >> gpurun_out/summary.txt; wc -l gpurun_out/launches_r1f.csv
cat gpurun_out/summary.txt
