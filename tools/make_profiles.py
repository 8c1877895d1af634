#!/usr/bin/env python3
"""Distil gpurun_out/ ncu artefacts into the tracked profiles/ directory.

usage: python tools/make_profiles.py <tag> [round]
  <tag>_launches.csv     -> profiles/<round>_launches_bench.csv (+ per-kernel share table)
  <tag>_top.ncu-rep      -> profiles/<round>_top_kernel.md / .json (ncu --page raw / source summaries)
and profiles/roofline_traffic.json (DRAM bytes per launch of the dominant kernel) read by bench.py.
"""
import collections, csv, gzip, io, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r1"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(',', '')); u = r[ui]
        v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
        k = re.sub(r'\(.*', '', r[ki])[:90]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    return agg

out = []
lp = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(lp):
    shutil.copy(lp, os.path.join(P, f"{rnd}_launches_bench.csv"))
    agg = launches(lp); tot = sum(v[1] for v in agg.values())
    out.append(f"## Launch list of `bench.py` (ncu --metrics gpu__time_duration.sum --clock-control none)\n")
    out.append(f"total kernel time {tot:.1f} us over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare SHARES)\n")
    out.append("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {t:.1f} | {100*t/tot:.1f}% | {t/n:.1f} |")
    out.append("")

rep = os.path.join(G, f"{tag}_top.ncu-rep")
traffic = {}
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
    summ = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for w in want:
            if w in hdr:
                d[w] = (r[hdr.index(w)], units[hdr.index(w)])
        summ.append(d)
        def num(key):
            v, u = d.get(key, ("0", ""))
            x = float(v.replace(",", ""))
            return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        kname = re.sub(r".*::", "", re.sub(r"<.*", "", d["kernel"].replace("void ", "")))
        traffic[kname] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    json.dump(summ, open(os.path.join(P, f"{rnd}_top_kernel_raw.json"), "w"), indent=1)
    out.append(f"## Top kernel, `ncu --set full --clock-control none --import-source on` ({os.path.basename(rep)})\n")
    for d in summ:
        out.append(f"**{d['kernel'][:110]}**\n")
        out.append("| metric | value |\n|---|---|")
        for w in want:
            if w in d: out.append(f"| {w} | {d[w][0]} {d[w][1]} |")
        out.append("")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    tmp = os.path.join(G, f"{tag}_top_source.csv.gz")
    with gzip.open(tmp, "wt") as fh: fh.write(src)
    s = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_source_summary.py"), tmp], capture_output=True, text=True).stdout
    out.append("### Source-level stall sampling / instruction mix\n\n```\n" + s + "```\n")
    json.dump(traffic, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)

open(os.path.join(P, f"{rnd}_bench_profile.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
