#!/bin/bash
# round 2, run 2: all GPU tests (full-depth / 257^3 / 513^3 / stage-3 twin), bench lines (configs[1] and [3]), ncu evidence of
# the SHIPPED kernels: launch list of one bench step, DRAM traffic of the DiT forward's GEMMs, --set full of attention / GEMM / decode
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x -s > $O/r2_02_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" $O/r2_02_pytest.log | tail -3
grep -E "full-depth DiT|latents after 10|257\^3 decode|513\^3:" $O/r2_02_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > $O/r2_02_bench.json 2> $O/r2_02_bench.err; echo "bench rc=$?"; cut -c1-1500 $O/r2_02_bench.json; tail -3 $O/r2_02_bench.err
timeout 400 python bench.py --workload vggt --steps 5 --warmup 3 > $O/r2_02_bench_vggt.json 2> $O/r2_02_bench_vggt.err; echo "vggt rc=$?"; cut -c1-1500 $O/r2_02_bench_vggt.json; tail -3 $O/r2_02_bench_vggt.err
# DRAM traffic + durations of every GEMM launch of one DiT forward (second forward: -s skips the first forward's 195 launches)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:linear_kernel -s 195 -c 195 --csv --log-file $O/r2_02_ncu_dit_gemm_traffic.csv python tools/prof_dit_forward.py > $O/r2_02_ncu_gemm.log 2>&1; echo "ncu gemm traffic rc=$?"
# --set full: the shipped attention kernel at the DiT shape, the CTA-pair GEMM (single-block linear1), attention at the decode shape
timeout 500 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 50 -c 1 -f -o $O/r2_02_prof_attn_dit python tools/prof_dit_forward.py > $O/r2_02_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:linear_kernel_2cta -s 300 -c 2 -f -o $O/r2_02_prof_gemm_dit python tools/prof_dit_forward.py > $O/r2_02_ncu_gemm2.log 2>&1; echo "ncu gemm rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"attention_kernel|lnpost_dot|layernorm_kernel" -s 12 -c 4 -f -o $O/r2_02_prof_decode python tools/prof_decode_chunk.py > $O/r2_02_ncu_dec.log 2>&1; echo "ncu decode rc=$?"
ls -la $O | tail -15
