#!/bin/bash
# ncu --set full of attention v4 (source-level stall sampling)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
R3G_ATTN=4 timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_attn4.log 2>&1 || { echo canary failed; exit 1; }
R3G_ATTN=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention -s 2 -c 1 -f -o gpurun_out/prof_attn_v4 python tools/prof_attn.py > gpurun_out/ncu_attn_v4.log 2>&1
tail -3 gpurun_out/ncu_attn_v4.log; ls -la gpurun_out/prof_attn_v4.ncu-rep
