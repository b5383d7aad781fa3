#!/bin/bash
# the full-object ncu launch list of the bench command (one object, 50 DiT steps), warm caches, no clock control
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout -s KILL 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo canary failed; exit 1; }
timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 30000 --csv --log-file gpurun_out/launches_r1f_full.csv python bench.py --steps 1 --warmup 0 --profile-mode > gpurun_out/ncu_launchlist_full.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches_r1f_full.csv; tail -1 gpurun_out/ncu_launchlist_full.log | cut -c1-300
gzip -kf gpurun_out/launches_r1f_full.csv; ls -la gpurun_out/launches_r1f_full.csv.gz
