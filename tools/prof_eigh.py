#!/usr/bin/env python3
"""lq_eigh_dev on the cfg5 Gram matrix (128 x 128): time + accuracy; LINALG_B200_JACOBI_TWO_SIDED=1 selects the old kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
ctx = nat.Context(0)
for n, rows in ((128, 1 << 20), (128, 4096), (64, 4096), (100, 300), (128, 100)):
    M0 = np.random.default_rng(9).standard_normal((rows, n)); G0 = M0.T @ M0
    dG, dl, dV = ctx.upload(G0), ctx.alloc(8 * n), ctx.alloc(8 * n * n)
    ms = []
    for _ in range(4):
        ctx.record(0); ctx.call("lq_eigh_dev", dG.ptr, n, dl.ptr, dV.ptr); ctx.record(1)
        ms.append(ctx.elapsed_ms(0, 1))
    lam = ctx.download(dl, (n,)); V = ctx.download(dV, (n, n))
    ref = np.linalg.eigvalsh(G0)[::-1]
    print(f"n={n} rows={rows}: {min(ms):.3f} ms  |lam-ref|/lmax {np.max(np.abs(lam-ref))/ref[0]:.2e}  orth {np.abs(V.T@V-np.eye(n)).max():.2e}  resid {np.abs(G0@V-V*lam).max()/ref[0]:.2e}", flush=True)
