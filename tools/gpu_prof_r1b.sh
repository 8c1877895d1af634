#!/bin/bash
# second batch of round-1 evidence: launch list of the blocked QR with the panel-wise look-ahead, --set full captures
# of the st.async panel kernel, the one-sided Jacobi kernel and the rank-128 update GEMM (TMA reduce-add epilogue)
mkdir -p gpurun_out
bash tools/ncu_launches.sh blocked8192g python tools/prof_blocked.py 8192 1 > gpurun_out/blocked8192g_summary.txt 2>&1
tail -14 gpurun_out/blocked8192g_summary.txt
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__cluster_size,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"
ncu --metrics $M --clock-control none -k regex:panel2_cluster -s 40 -c 2 --csv --log-file gpurun_out/panel2_ncu.csv python tools/prof_blocked.py 4096 1 > gpurun_out/panel2_ncu.log 2>&1
ncu --metrics $M --clock-control none -k regex:jacobi1s -c 3 --csv --log-file gpurun_out/jacobi1s_ncu.csv python tools/prof_eigh.py > gpurun_out/jacobi1s_ncu.log 2>&1
ncu --metrics $M --clock-control none -k regex:gemm_ -c 6 --csv --log-file gpurun_out/gemm_red_ncu.csv python tools/prof_gemm.py > gpurun_out/gemm_red_ncu.log 2>&1
ls -la gpurun_out/*_ncu.csv
