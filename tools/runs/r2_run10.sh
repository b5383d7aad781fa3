#!/bin/bash
# round 2, run 10: full GPU suite with both VGGT heads on r3g kernels; config-4 bench line; final single-GPU line
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -s > $O/r2_10_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $O/r2_10_pytest.log | tail -3; grep -E "DPT on r3g|^E  |Error" $O/r2_10_pytest.log | head -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_10_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2_10_smoke.log
timeout 400 python bench.py --workload vggt --steps 5 --warmup 3 > $O/r2_10_bench_vggt.json 2> $O/r2_10_bench_vggt.err; echo "vggt rc=$?"; python -c "
import json; d=json.load(open('$O/r2_10_bench_vggt.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['aggregator']['ms'], d['gpu_launches'], d.get('cpu_baseline',{}).get('value'))"; tail -3 $O/r2_10_bench_vggt.err
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2_10_bench.json 2> $O/r2_10_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/r2_10_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['stages_ms_last_object'], d['clocks'], d['cpu_baseline']['value'])"; tail -3 $O/r2_10_bench.err
