#!/bin/bash
# round 2, run 6 (2 GPUs): the streaming mesh gather over NCCL -- bench line at N=2 and the stage-3 twin under torchrun
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python -m pytest tests/test_gpu_postprocess.py tests/test_gpu_vggt.py -q -m gpu --timeout 200 > $O/r2_06_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_06_pytest.log
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 2 > $O/r2_06_bench_2gpu.json 2> $O/r2_06_bench_2gpu.err; echo "bench2 rc=$?"; python -c "
import json; d=json.loads(open('$O/r2_06_bench_2gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['clocks'])"; tail -5 $O/r2_06_bench_2gpu.err
# stage-3 twin, 2 ranks, 5 crops (one skipped), few steps / small grid
python - <<'PY'
import os, sys, yaml
sys.path.insert(0, '.')
from bench import synthetic_crop
os.makedirs('/tmp/s3/in', exist_ok=True); os.makedirs('/tmp/s3/out', exist_ok=True)
for n, s in (('a__1.png', 1), ('b__2.png', 2), ('c__3.png', 3), ('d__4.png', 4), ('wall__0.png', 5)):
    synthetic_crop(s).save('/tmp/s3/in/' + n)
yaml.safe_dump(dict(use_banana=False, input_folder_hy='/tmp/s3/in', output_folder_hy='/tmp/s3/out', num_inf_steps_hy=3,
                    octree_resolution_hy=96, num_chunks_hy=16000, seed=1234567, mini=False), open('/tmp/s3/config.yaml', 'w'))
PY
timeout 900 $TR stages/2d_to_3d_models/run.py --config /tmp/s3/config.yaml --random-weights > $O/r2_06_stage3_2gpu.log 2>&1; echo "stage3 rc=$?"; grep -E "vertices|wrote|Error|error" $O/r2_06_stage3_2gpu.log | tail -8; ls -la /tmp/s3/out/*/ | head -12
timeout 600 $TR bench.py --gpus 2 --objects 4 --octree 512 --warmup 1 --no-cpu-baseline > $O/r2_06_bench_2gpu_512.json 2> $O/r2_06_bench_2gpu_512.err; echo "bench 512 rc=$?"; python -c "
import json; d=json.loads(open('$O/r2_06_bench_2gpu_512.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['scaling'], d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms_last_object'], d['config']['workload'])"; tail -5 $O/r2_06_bench_2gpu_512.err
