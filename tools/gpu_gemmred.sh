#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_red.log 2>&1; echo "gemm rc=$?"; cat gpurun_out/gemm_red.log
for v in 1 0; do
  LINALG_B200_PANEL=$v timeout 300 python tools/blocked_bench.py 2048,8192 > gpurun_out/blocked_red_p$v.log 2>&1; echo "blocked v$v rc=$?"
  tail -6 gpurun_out/blocked_red_p$v.log
done
