#!/bin/bash
# round 2, run 18: the whole GPU suite on the final tree (after the vertex-pass change of run 17)
mkdir -p gpurun_out
timeout 110 python -m pytest tests -q -m gpu -x --timeout 100 > gpurun_out/r2_18_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_18_pytest.log; grep -E "^E |FAILED" gpurun_out/r2_18_pytest.log | head
