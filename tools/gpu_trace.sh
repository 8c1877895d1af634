#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
  LINALG_B200_PANEL=$v timeout 200 python tools/blocked_trace.py 8192 2> gpurun_out/blocked_trace_p$v.log; echo "trace v$v rc=$?"
  LINALG_B200_PANEL=$v timeout 200 python tools/blocked_split.py 8192 2>&1 | tee gpurun_out/blocked_split_p$v.log
done
LINALG_B200_NO_LOOKAHEAD=1 timeout 200 python tools/blocked_split.py 8192 2>&1 | tee gpurun_out/blocked_split_nola.log
