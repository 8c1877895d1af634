#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/blocked_bench.py 2048,4096,8192 > gpurun_out/blocked_pw.log 2>&1; echo "pw rc=$?"; tail -8 gpurun_out/blocked_pw.log
LINALG_B200_NO_PANELWISE=1 timeout 300 python tools/blocked_bench.py 2048,4096,8192 > gpurun_out/blocked_nopw.log 2>&1; echo "nopw rc=$?"; tail -5 gpurun_out/blocked_nopw.log
timeout 200 python tools/blocked_trace.py 8192 2> gpurun_out/blocked_trace_pw.log; echo "trace rc=$?"
