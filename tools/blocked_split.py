#!/usr/bin/env python3
"""Split of the blocked QR time: factorisation only (least-squares entry, no Q) vs factorisation + Q."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
ctx = nat.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
A = np.random.default_rng(5).standard_normal((n, n)); B = np.random.default_rng(6).standard_normal((n, 2))
dA, dB, dX = ctx.upload(A), ctx.upload(B), ctx.alloc(8 * n * 2)
dQ, dR = ctx.alloc(A.nbytes), ctx.alloc(A.nbytes)
for name, fn in (("factor+rhs (no Q)", lambda: ctx.call("lq_lstsq_householder_batched_dev", dA.ptr, dB.ptr, 1, n, n, 2, dX.ptr)),
                 ("factor + Q", lambda: ctx.call("lq_householder_qr_dev", dA.ptr, n, n, dQ.ptr, dR.ptr))):
    ms = []
    for _ in range(4):
        ctx.record(0); fn(); ctx.record(1); ms.append(ctx.elapsed_ms(0, 1))
    print(f"{name}: {min(ms[1:]):.2f} ms", flush=True)
