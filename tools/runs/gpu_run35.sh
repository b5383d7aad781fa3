#!/bin/bash
# last GPU seconds of the round: the two-threads-per-row attention family with the packed-fp32 softmax (family 3, variants 2/3)
mkdir -p gpurun_out
for v in 2 3; do
  R3G_ATTN=3 R3G_ATTN_VARIANT=$v R3G_MB_ONLY=attention R3G_MB_OUT=mb_attn_a3v$v.json timeout -s KILL 40 python tools/microbench.py > gpurun_out/mb_attn_a3v$v.log 2>&1
  echo "attn=3 variant=$v: $(grep -o "'tflops': [0-9.]*" gpurun_out/mb_attn_a3v$v.log | tr '\n' ' ')"
done
R3G_ATTN=3 R3G_ATTN_VARIANT=2 timeout -s KILL 30 python -m pytest tests/test_gpu_linear_attention.py -q -m gpu --timeout 25 -k "test_attention and not kept" 2>&1 | tail -2
