#!/bin/bash
# usage: tools/ncu_launches.sh <tag> <cmd...>   -> gpurun_out/<tag>_launches.csv (per-launch gpu time)
TAG=$1; shift
mkdir -p gpurun_out
"$@" > gpurun_out/${TAG}_plain.log 2>&1 || { echo plain failed; tail gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_plain.log
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/${TAG}_launches.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
    k = r[ki][:70]; agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print('total kernel time %.1f us over %d launches' % (tot, sum(v[0] for v in agg.values())))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print('%8.1f us %5.1f%% %5d x %7.1f us  %s' % (t, 100*t/tot, n, t/n, k))
PY
