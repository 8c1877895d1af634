#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "8192 K=128|M=256 N=384|M=1048576" | tee gpurun_out/gemm_upd.log
LINALG_B200_NO_UPD_GEMM=1 timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "8192 K=128" 
timeout 300 python tools/blocked_bench.py 2048,4096,8192 > gpurun_out/blocked_upd.log 2>&1; echo "rc=$?"; tail -9 gpurun_out/blocked_upd.log
