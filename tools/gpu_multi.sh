#!/bin/bash
# usage: tools/gpu_multi.sh <ngpu> <tag>  (run on a multi-GPU box): 2-rank NCCL parity tests + the N-GPU bench with checked extras
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_multi_box.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rA > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/${TAG}_pytest_multi.log
tail -8 gpurun_out/${TAG}_pytest_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/${TAG}_bench_n${N}.json; tail -5 gpurun_out/${TAG}_bench_n${N}.err
