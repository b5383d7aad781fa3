#!/bin/bash
# round 2, run 15: marching cubes v3 (sign-bit volume, crossed-cell list, one thread per crossed cell): parity, probe,
# per-kernel durations of the probe under ncu
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_mc.py tests/test_gpu_postprocess.py tests/test_gpu_flashvdm.py -q -m gpu -x --timeout 300 > $O/r2_15_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2_15_pytest.log; grep -E "^E " $O/r2_15_pytest.log | head -10
timeout 300 python tools/mc_probe.py > $O/r2_15_mc_probe.json 2> $O/r2_15_mc_probe.err; echo "probe rc=$?"; tail -8 $O/r2_15_mc_probe.err; grep "r3g mc" $O/r2_15_mc_probe.json
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:mc_ --csv --log-file $O/r2_15_mc_launches.csv python tools/mc_probe.py > /dev/null 2> $O/r2_15_ncu.err; echo "ncu rc=$?"; wc -l $O/r2_15_mc_launches.csv
