#!/bin/bash
# round 2, run 13: marching cubes with the ballot classify pass / info words / parallel scan: parity tests, then the probe
# on the new library and on the previous one (tools/runs/_prev/libr3g_prev.so, built from the parent commit's mc.cu)
mkdir -p gpurun_out
O=gpurun_out
L=3d-re-gen_b200/r3g/libr3g.so
timeout 600 python -m pytest tests/test_gpu_mc.py tests/test_gpu_postprocess.py tests/test_gpu_flashvdm.py -q -m gpu -x --timeout 300 > $O/r2_13_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2_13_pytest.log; grep -E "^E " $O/r2_13_pytest.log | head -10
timeout 300 python tools/mc_probe.py > $O/r2_13_mc_probe.json 2> $O/r2_13_mc_probe.err; echo "probe rc=$?"; tail -8 $O/r2_13_mc_probe.err; grep "r3g mc" $O/r2_13_mc_probe.json
if [ -f tools/runs/_prev/libr3g_prev.so ]; then
  cp $L /tmp/libr3g_new.so; cp tools/runs/_prev/libr3g_prev.so $L
  timeout 300 python tools/mc_probe.py > $O/r2_13_mc_probe_prev.json 2> $O/r2_13_mc_probe_prev.err; echo "probe(prev) rc=$?"; tail -8 $O/r2_13_mc_probe_prev.err; grep "r3g mc" $O/r2_13_mc_probe_prev.json
  cp /tmp/libr3g_new.so $L
fi
