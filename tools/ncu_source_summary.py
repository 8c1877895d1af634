#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` dump: stall reasons and instruction mix per kernel."""
import csv, gzip, sys, collections, re
path = sys.argv[1]
op = gzip.open if path.endswith('.gz') else open
with op(path, 'rt') as fh:
    rd = csv.reader(fh)
    kern = None; hdr = None; data = {}
    for row in rd:
        if not row: continue
        if row[0] == 'Kernel Name':
            kern = row[1]; data[kern] = []; hdr = None; continue
        if row[0] == 'Address':
            hdr = row; continue
        if hdr and kern: data[kern].append(dict(zip(hdr, row)))
for kern, rows in data.items():
    print('=' * 100); print(kern[:120])
    stalls = collections.Counter(); ops = collections.Counter(); samples = 0
    opsamp = collections.Counter()
    for r in rows:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r['Source'])
        opc = m.group(2).split('.')[0] if m else '?'
        n = int(r['Instructions Executed'] or 0)
        ops[opc] += n
        s = int(r['# Samples'] or 0); samples += s; opsamp[opc] += s
        for k, v in r.items():
            if k.startswith('stall_') and 'Not Issued' not in k and v not in ('', '-'):
                stalls[k] += int(v)
    tot = sum(ops.values())
    print('instructions executed (warp-level): %d' % tot)
    print('  mix:', ', '.join('%s %.1f%%' % (k, 100.0 * v / tot) for k, v in ops.most_common(14)))
    print('samples: %d' % samples)
    print('  stalls:', ', '.join('%s %.1f%%' % (k[6:], 100.0 * v / max(1, samples)) for k, v in stalls.most_common(10)))
    print('  samples by opcode:', ', '.join('%s %.1f%%' % (k, 100.0 * v / max(1, samples)) for k, v in opsamp.most_common(10)))
    exc = sum(int(r['L1 Wavefronts Shared Excessive'] or 0) for r in rows); wf = sum(int(r['L1 Wavefronts Shared'] or 0) for r in rows)
    print('  shared wavefronts %d, excessive %d' % (wf, exc))
