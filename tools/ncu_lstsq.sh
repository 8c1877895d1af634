#!/bin/bash
# usage: tools/ncu_lstsq.sh <tag> [nsys]   (run on the GPU box): ncu --set full of the cfg3 least-squares kernel
set -u
TAG=$1; NS=${2:-16384}
mkdir -p gpurun_out
python tools/prof_lstsq.py $NS > gpurun_out/${TAG}_plain.log 2>&1 || { echo plain run failed; tail gpurun_out/${TAG}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:lstsq -s 2 -c 1 -o /tmp/${TAG} python tools/prof_lstsq.py $NS > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_source.csv 2>/dev/null
ncu -i /tmp/${TAG}.ncu-rep --page source --csv --print-source cuda > gpurun_out/${TAG}_cuda.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_source.csv gpurun_out/${TAG}_cuda.csv
ls -la /tmp/${TAG}.ncu-rep
tail -3 gpurun_out/${TAG}_plain.log
