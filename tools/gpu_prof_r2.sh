#!/bin/bash
# round-2 evidence (run on the GPU box): launch list of bench.py + --set full of the headline kernel at 2^20, --set full of the
# K3 kernel, metric captures of the FP64 probes / SYRK / tall-skinny pipeline, launch list of one SVD + TSQR call
mkdir -p gpurun_out
bash tools/ncu_bench.sh r2 hh_qr32_c8 > gpurun_out/r2_ncu_bench.log 2>&1; tail -3 gpurun_out/r2_ncu_bench.log
bash tools/ncu_lstsq.sh r2_lstsq_final > gpurun_out/r2_ncu_lstsq.log 2>&1; tail -2 gpurun_out/r2_ncu_lstsq.log
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_fp64.sum,sm__cycles_elapsed.avg.per_second"
python tools/prof_ts_r2.py > gpurun_out/r2_ts_plain.log 2>&1 || { echo plain failed; tail gpurun_out/r2_ts_plain.log; }
ncu --metrics $M --clock-control none -k regex:'probe|syrk' -c 12 --csv --log-file gpurun_out/r2_probe_syrk_ncu.csv python tools/prof_ts_r2.py > gpurun_out/r2_probe_syrk_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ts_launches.csv python tools/prof_ts_r2.py > gpurun_out/r2_ts_launches.log 2>&1
tail -8 gpurun_out/r2_ts_plain.log; ls -la gpurun_out | tail -12
