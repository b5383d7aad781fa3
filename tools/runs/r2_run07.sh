#!/bin/bash
# round 2, run 7: full GPU suite (postprocess, aggregator graph, conditioner), smoke, final single-GPU bench lines
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x > $O/r2_07_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2_07_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_07_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2_07_smoke.log
timeout 400 python bench.py --workload vggt --steps 5 --warmup 3 > $O/r2_07_bench_vggt.json 2> $O/r2_07_bench_vggt.err; echo "vggt rc=$?"; python -c "
import json; d=json.load(open('$O/r2_07_bench_vggt.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['aggregator'], d.get('cpu_baseline',{}).get('value'))"; tail -3 $O/r2_07_bench_vggt.err
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2_07_bench.json 2> $O/r2_07_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/r2_07_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['stages_ms_last_object'], d['clocks'], d['cpu_baseline']['value'], d['gpu_launches'])"; tail -3 $O/r2_07_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_07_bench_reference.json 2> $O/r2_07_bench_reference.err; echo "ref rc=$?"; cut -c1-300 $O/r2_07_bench_reference.json
timeout 600 python bench.py --octree 512 --steps 2 --warmup 1 --no-cpu-baseline > $O/r2_07_bench_512.json 2> $O/r2_07_bench_512.err; echo "512 rc=$?"; python -c "
import json; d=json.load(open('$O/r2_07_bench_512.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms_last_object'], d['object'])"; tail -3 $O/r2_07_bench_512.err
for what in attention linear; do R3G_MB_ONLY=$what R3G_MB_OUT=r2_07_mb_$what.json timeout 300 python tools/microbench.py > $O/r2_07_mb_$what.log 2>&1; echo "mb $what rc=$?"; done
