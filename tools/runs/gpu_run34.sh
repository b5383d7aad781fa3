#!/bin/bash
# 2-GPU sanity run of the bench contract after the r1f bench edits (1 warm-up + 1 timed object per rank)
mkdir -p gpurun_out
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_2gpu_r1f.json 2> gpurun_out/bench_2gpu_r1f.err
echo "exit $?"; cut -c1-1500 gpurun_out/bench_2gpu_r1f.json; tail -3 gpurun_out/bench_2gpu_r1f.err
