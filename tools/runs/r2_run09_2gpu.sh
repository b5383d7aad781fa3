#!/bin/bash
# round 2, run 9 (2 GPUs): re-check after the union-find flatten fix and the gather-capacity agreement; camera head on r3g kernels
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 400 python -m pytest tests/test_gpu_postprocess.py tests/test_gpu_vggt.py -q -m gpu --timeout 200 > $O/r2_09_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2_09_pytest.log; grep -E "^E " $O/r2_09_pytest.log | head -10
timeout 300 python bench.py --workload vggt --steps 5 --warmup 3 --no-cpu-baseline > $O/r2_09_bench_vggt.json 2> $O/r2_09_bench_vggt.err; echo "vggt rc=$?"; python -c "
import json; d=json.load(open('$O/r2_09_bench_vggt.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['aggregator']['ms'])"; tail -3 $O/r2_09_bench_vggt.err
timeout 700 $TR bench.py --gpus 2 --objects 4 --octree 512 --warmup 1 --no-cpu-baseline > $O/r2_09_bench_2gpu_512.json 2> $O/r2_09_bench_2gpu_512.err; echo "bench 512 rc=$?"; python -c "
import json; d=json.loads(open('$O/r2_09_bench_2gpu_512.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['scaling'], d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms_last_object'], d['config']['workload'])"; grep -E "Error|error" $O/r2_09_bench_2gpu_512.err | head -5
