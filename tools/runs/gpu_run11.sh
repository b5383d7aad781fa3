#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for d in 0 1 2 4 3 7; do
  echo "=== R3G_ATTN=2 DBG=$d"
  R3G_ATTN=2 R3G_ATTN_DBG=$d R3G_MB_ONLY=attention R3G_MB_OUT=mb_attn_dbg$d.json timeout 120 python tools/microbench.py 2>&1 | cut -c1-120 | head -1
done
