#!/bin/bash
# round 2, run 4: FlashVDM decoder vs the reference fixture, whole-decode CUDA graph, full GPU suite, bench
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_flashvdm.py -q -m gpu --timeout 300 -s > $O/r2_04_flashvdm.log 2>&1; echo "flashvdm rc=$?"; grep -E "two levels|dense level|FlashVDM 252|passed|failed|Error|error" $O/r2_04_flashvdm.log | head -20
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x > $O/r2_04_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2_04_pytest.log
R3G_PROBE_OUT=r2_04_decode_probe.json timeout 300 python tools/decode_probe.py > $O/r2_04_decode_probe.log 2>&1; echo "probe rc=$?"; grep -E "rep" $O/r2_04_decode_probe.log | cut -c1-200
timeout 600 python bench.py --steps 3 --warmup 3 > $O/r2_04_bench.json 2> $O/r2_04_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/r2_04_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['launches_timed'], d['stages_ms_last_object'], d['clocks'], d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'))"; tail -3 $O/r2_04_bench.err
