#!/bin/bash
# End-of-round session on one GPU: smoke, GPU parity tests, bench (N=1), reference arm, the reference-style benchmark table.
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv; nproc; free -g | head -2; } > gpurun_out/r2_box.txt 2>&1
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_gpu.log
tail -6 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
timeout 600 python tools/benchmark_qr_table.py > gpurun_out/r2_benchmark_qr_table.md 2>&1; echo "table rc=$?"; tail -12 gpurun_out/r2_benchmark_qr_table.md
timeout 300 python tools/fuzz_shapes.py > gpurun_out/r2_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -3 gpurun_out/r2_fuzz.log
