#!/bin/bash
mkdir -p gpurun_out
timeout 180 python tools/panel_bench.py > gpurun_out/panel_bench.log 2>&1; echo "panel rc=$?"
cat gpurun_out/panel_bench.log
for v in 1 2 0; do
  LINALG_B200_PANEL=$v timeout 300 python tools/blocked_bench.py 2048,8192 > gpurun_out/blocked_p$v.log 2>&1; echo "blocked v$v rc=$?"
  tail -6 gpurun_out/blocked_p$v.log
done
