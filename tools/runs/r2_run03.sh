#!/bin/bash
# round 2, run 3: grouped img+txt GEMM launches and programmatic dependent launch: correctness, then the DiT step in-graph
# with each feature on / off; where the decode's time goes (host enqueue vs device, per kernel)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_linear_pair.py tests/test_gpu_linear_attention.py tests/test_gpu_models.py tests/test_gpu_pipeline.py tests/test_gpu_rowops.py -q -m gpu --timeout 300 -x > $O/r2_03_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2_03_pytest.log
for pdl in 0 1; do for grp in 0 1; do
  R3G_PDL=$pdl R3G_GROUP=$grp R3G_ABLATE_OUT=r2_03_ablate_pdl${pdl}_grp${grp}.json timeout 300 python tools/ablate_dit.py > $O/r2_03_ablate_pdl${pdl}_grp${grp}.log 2>&1
  echo "pdl=$pdl group=$grp rc=$? $(grep -o '"full_ms": [0-9.]*' $O/r2_03_ablate_pdl${pdl}_grp${grp}.log | tail -1)"; grep -E "^linear|^attention|^layernorm|^gemv" $O/r2_03_ablate_pdl${pdl}_grp${grp}.log | cut -c1-150
done; done
R3G_PDL=0 R3G_PROBE_OUT=r2_03_decode_probe_pdl0.json timeout 300 python tools/decode_probe.py > $O/r2_03_decode_probe_pdl0.log 2>&1; echo "probe pdl0 rc=$?"; grep -E "rep|chunk_graph|^k" $O/r2_03_decode_probe_pdl0.log | cut -c1-200
R3G_PDL=1 R3G_PROBE_OUT=r2_03_decode_probe_pdl1.json timeout 300 python tools/decode_probe.py > $O/r2_03_decode_probe_pdl1.log 2>&1; echo "probe pdl1 rc=$?"; grep -E "rep|chunk_graph" $O/r2_03_decode_probe_pdl1.log | cut -c1-200
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_03_bench.json 2> $O/r2_03_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/r2_03_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['launches_timed'], d['stages_ms_last_object'], d['clocks'])"
