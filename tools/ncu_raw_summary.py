#!/usr/bin/env python3
"""Print the handful of ncu raw-page metrics the notes quote, one block per profiled kernel."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d['Kernel Name'][:90])
    for w in want:
        print('   %-75s %s' % (w, d.get(w)))
    st = sorted(((float(d[h]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]) for h in stall if d.get(h)), reverse=True)
    print('   stalls per issue:', ', '.join('%s %.2f' % (n, v) for v, n in st[:9]))
