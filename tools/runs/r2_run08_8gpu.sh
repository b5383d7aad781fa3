#!/bin/bash
# round 2, run 8 (8 GPUs): BASELINE configs[2] (8 objects of one scene, one per GPU), configs[4] (32 crops, 512^3) and a short weak-scaling line
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
show() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d['n_gpus'], d['scaling'], 'value', round(d['value'],4), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],4), d['stages_ms_last_object'], d['clocks'])" $1; }
timeout 900 $TR bench.py --gpus 8 --steps 3 --warmup 2 > $O/r2_08_bench_8gpu_weak.json 2> $O/r2_08_bench_8gpu_weak.err; echo "weak rc=$?"; show $O/r2_08_bench_8gpu_weak.json; tail -3 $O/r2_08_bench_8gpu_weak.err
timeout 900 $TR bench.py --gpus 8 --objects 8 --crops 2400 --warmup 2 > $O/r2_08_bench_8gpu_config3.json 2> $O/r2_08_bench_8gpu_config3.err; echo "config3 rc=$?"; show $O/r2_08_bench_8gpu_config3.json; tail -3 $O/r2_08_bench_8gpu_config3.err
timeout 1500 $TR bench.py --gpus 8 --objects 32 --octree 512 --warmup 1 > $O/r2_08_bench_8gpu_config5.json 2> $O/r2_08_bench_8gpu_config5.err; echo "config5 rc=$?"; show $O/r2_08_bench_8gpu_config5.json; tail -3 $O/r2_08_bench_8gpu_config5.err
