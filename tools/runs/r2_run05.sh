#!/bin/bash
# round 2, run 5: conditioner on r3g kernels, run_VGGT composite; ncu: DRAM traffic of the DiT forward's GEMM launches after
# grouping; launch list of ONE timed object of bench.py (shares of the step)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_vggt.py tests/test_gpu_flashvdm.py tests/test_gpu_pipeline.py -q -m gpu --timeout 300 -s > $O/r2_05_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FlashVDM 252|Error" $O/r2_05_pytest.log | tail -6
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:linear_kernel -s 131 -c 131 --csv --log-file $O/r2_05_ncu_dit_gemm_traffic.csv python tools/prof_dit_forward.py > $O/r2_05_ncu_gemm.log 2>&1; echo "ncu gemm traffic rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:linear_kernel_2cta -s 140 -c 3 -f -o $O/r2_05_prof_gemm_dit python tools/prof_dit_forward.py > $O/r2_05_ncu_gemm2.log 2>&1; echo "ncu gemm full rc=$?"
timeout 2400 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_05_launches_object.csv python bench.py --steps 1 --warmup 1 --profile-mode > $O/r2_05_launches.log 2>&1; echo "ncu launch list rc=$?"; tail -2 $O/r2_05_launches.log | cut -c1-300; wc -l $O/r2_05_launches_object.csv
