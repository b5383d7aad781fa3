#!/bin/bash
# round 2, run 11: where does the e2e step spend its extra time (host breakdown)?  + the image-processor fix
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2_11_bench.json 2> $O/r2_11_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/r2_11_bench.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['stages_ms_last_object'], d['clocks'])"; tail -3 $O/r2_11_bench.err
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_full_config.py -q -m gpu --timeout 400 -x > $O/r2_11_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_11_pytest.log
