#!/bin/bash
# round 2, run 12 (8 GPUs): the N > 1 bench path after the change to sequential arms + one gather per arm
# (receive buffers, pinned ring and NCCL point-to-point channels set up before the timed region)
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
show() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d['n_gpus'], d['scaling'], 'value', round(d['value'],4), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],4), d['e2e']['ms_per_step'], d['stages_ms_last_object'], d['clocks'])" $1; }
timeout 240 $TR bench.py --gpus 8 --steps 3 --warmup 3 > $O/r2_12_bench_8gpu_weak.json 2> $O/r2_12_bench_8gpu_weak.err; echo "weak rc=$?"; show $O/r2_12_bench_8gpu_weak.json; tail -3 $O/r2_12_bench_8gpu_weak.err
timeout 200 $TR bench.py --gpus 8 --objects 8 --crops 2400 --warmup 3 > $O/r2_12_bench_8gpu_config3.json 2> $O/r2_12_bench_8gpu_config3.err; echo "config3 rc=$?"; show $O/r2_12_bench_8gpu_config3.json; tail -3 $O/r2_12_bench_8gpu_config3.err
