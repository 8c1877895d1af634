#!/usr/bin/env python3
"""Per-outer-block timeline of the blocked QR factorisation (LINALG_B200_TRACE_BLOCKS=1 prints it to stderr)."""
import os, sys
os.environ["LINALG_B200_TRACE_BLOCKS"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
ctx = nat.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
A = np.random.default_rng(5).standard_normal((n, n)); B = np.random.default_rng(6).standard_normal((n, 2))
dA, dB, dX = ctx.upload(A), ctx.upload(B), ctx.alloc(8 * n * 2)
for it in range(2):
    print(f"# pass {it}", file=sys.stderr, flush=True)
    ctx.call("lq_lstsq_householder_batched_dev", dA.ptr, dB.ptr, 1, n, n, 2, dX.ptr)
    ctx.sync()
