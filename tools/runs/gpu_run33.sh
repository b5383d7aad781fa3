#!/bin/bash
# final build check (after the helper de-duplication): canary + the kernel-level parity tests
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout -s KILL 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo canary failed; tail -5 gpurun_out/smoke.log; exit 1; }
tail -1 gpurun_out/smoke.log
timeout -s KILL 200 python -m pytest tests/test_gpu_rowops.py tests/test_gpu_linear_attention.py tests/test_gpu_models.py -q -m gpu --timeout 100 -k "not kept_generations" 2>&1 | tail -3
