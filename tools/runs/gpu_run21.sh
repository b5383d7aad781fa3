#!/bin/bash
# row-kernel micro-benchmarks + a short ncu launch list (2 DiT steps + full decode) with the current kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo canary failed; exit 1; }
R3G_MB_ONLY=rowops R3G_MB_OUT=mb_rowops.json timeout 200 python tools/microbench.py > gpurun_out/mb_rowops.log 2>&1; cut -c1-150 gpurun_out/mb_rowops.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_short.csv python bench.py --steps 1 --warmup 0 --profile-mode --dit-steps 2 > gpurun_out/ncu_short.log 2>&1
tail -2 gpurun_out/ncu_short.log | cut -c1-300; wc -l gpurun_out/launches_short.csv
