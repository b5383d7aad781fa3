#!/bin/bash
# round 2, run 17: vertex pass with the owned-edge list (float64 loop runs nv times); parity, probe, and one
# `ncu --set full` capture of a complete marching-cubes call on the dense 513^3 volume (8 kernels)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_mc.py tests/test_gpu_postprocess.py tests/test_gpu_flashvdm.py tests/test_gpu_pipeline.py -q -m gpu -x --timeout 300 > $O/r2_17_pytest_mc.log 2>&1; rc=$?; echo "pytest mc rc=$rc"; tail -3 $O/r2_17_pytest_mc.log; grep -E "^E " $O/r2_17_pytest_mc.log | head -10
[ $rc -ne 0 ] && exit 1
timeout 300 python tools/mc_probe.py > $O/r2_17_mc_probe.json 2> $O/r2_17_mc_probe.err; echo "probe rc=$?"; tail -8 $O/r2_17_mc_probe.err; grep "r3g mc" $O/r2_17_mc_probe.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mc_ -s 216 -c 8 -f -o $O/r2_17_prof_mc_513_dense python tools/mc_probe.py > /dev/null 2> $O/r2_17_ncu.err; echo "ncu rc=$?"; ls -la $O/r2_17_prof_mc_513_dense.ncu-rep
