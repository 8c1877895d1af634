#!/usr/bin/env python3
"""Where a 256^2 (or n^2) householder_qr call spends its time: factorisation timeline (LINALG_B200_TRACE_BLOCKS) + total."""
import os, sys
os.environ["LINALG_B200_TRACE_BLOCKS"] = "1"
os.environ["LINALG_B200_NO_GRAPH"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
ctx = nat.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
A = np.random.default_rng(5).standard_normal((n, n))
dA, dQ, dR = ctx.upload(A), ctx.alloc(A.nbytes), ctx.alloc(A.nbytes)
for it in range(3):
    print(f"# pass {it}", file=sys.stderr, flush=True)
    ctx.record(0); ctx.call("lq_householder_qr_dev", dA.ptr, n, n, dQ.ptr, dR.ptr); ctx.record(1)
    print(f"# total {ctx.elapsed_ms(0, 1):.3f} ms", file=sys.stderr, flush=True)
