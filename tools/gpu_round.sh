#!/bin/bash
# One GPU-box session: environment facts, smoke, GPU parity tests, bench (N=1), reference arm.
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv
  nproc; free -g | head -2
  cat /sys/fs/cgroup/memory.max 2>/dev/null
} > gpurun_out/box.txt 2>&1
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; echo "ref rc=$?"
