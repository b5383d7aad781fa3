#!/bin/bash
# round 2, run 14: marching cubes with persistent classify (next tile's points in flight), emit kernels over the list of
# emitting tiles: parity tests, probe
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_mc.py tests/test_gpu_postprocess.py tests/test_gpu_flashvdm.py -q -m gpu -x --timeout 300 > $O/r2_14_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2_14_pytest.log; grep -E "^E " $O/r2_14_pytest.log | head -10
timeout 300 python tools/mc_probe.py > $O/r2_14_mc_probe.json 2> $O/r2_14_mc_probe.err; echo "probe rc=$?"; tail -8 $O/r2_14_mc_probe.err; grep "r3g mc" $O/r2_14_mc_probe.json
