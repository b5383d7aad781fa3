#!/bin/bash
# round 2, run 16: marching cubes with the flat sign-bit stream (16-byte loads): parity first (stop if it fails), probe,
# per-kernel list under ncu, then the whole GPU suite, smoke and the default bench line on the final tree
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_mc.py tests/test_gpu_postprocess.py tests/test_gpu_flashvdm.py -q -m gpu -x --timeout 300 > $O/r2_16_pytest_mc.log 2>&1; rc=$?; echo "pytest mc rc=$rc"; tail -5 $O/r2_16_pytest_mc.log; grep -E "^E " $O/r2_16_pytest_mc.log | head -10
[ $rc -ne 0 ] && exit 1
timeout 300 python tools/mc_probe.py > $O/r2_16_mc_probe.json 2> $O/r2_16_mc_probe.err; echo "probe rc=$?"; tail -8 $O/r2_16_mc_probe.err; grep "r3g mc" $O/r2_16_mc_probe.json
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:mc_ --csv --log-file $O/r2_16_mc_launches.csv python tools/mc_probe.py > /dev/null 2> $O/r2_16_ncu.err; echo "ncu rc=$?"; wc -l $O/r2_16_mc_launches.csv
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > $O/r2_16_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2_16_pytest.log; grep -E "^E |FAILED" $O/r2_16_pytest.log | head -10
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_16_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2_16_smoke.log
timeout 600 python bench.py > $O/r2_16_bench.json 2> $O/r2_16_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('$O/r2_16_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['stages_ms_last_object'], d['clocks'], d.get('cpu_baseline',{}).get('value'))"; tail -2 $O/r2_16_bench.err
