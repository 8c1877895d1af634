#!/usr/bin/env python3
"""Pipelined hh32 kernel (variant 13) against the one-shot kernel (variant 6: the same scalar chain) and the oracle, ragged batch sizes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
from oracle import linalg_oracle as orc
ctx = nat.Context(0)
rng = np.random.default_rng(7)
ok = True
for batch in (1, 2, 3, 5, 64, 1001, 4736, 4737, 20000):
    A = rng.standard_normal((batch, 32, 32))
    if batch >= 64:
        A[3, :, 5] = 0.0              # skipped reflector
        A[7, :, 9] = A[7, :, 2]       # dependent column
        A[11] = np.triu(A[11])        # already triangular
    dA = ctx.upload(A)
    outs = {}
    for v in (6, 13):
        dQ, dR = ctx.upload(np.full_like(A, np.nan)), ctx.upload(np.full_like(A, np.nan))
        ctx.call("lq_householder_qr_batched_dev", dA.ptr, batch, 32, 32, dQ.ptr, dR.ptr, v)
        outs[v] = (ctx.download(dQ, A.shape), ctx.download(dR, A.shape))
        dQ.free(); dR.free()
    dq = float(np.max(np.abs(outs[6][0] - outs[13][0]))); dr = float(np.max(np.abs(outs[6][1] - outs[13][1])))
    nb = min(batch, 48)
    Qo, Ro = orc.householder_qr_batched(A[:nb])
    eq, er = orc.rel_max_err(outs[13][0][:nb], Qo), orc.rel_max_err(outs[13][1][:nb], Ro)
    nan = bool(np.isnan(outs[13][0]).any() or np.isnan(outs[13][1]).any())
    good = dq == 0.0 and dr == 0.0 and eq < 1e-10 and er < 1e-10 and not nan
    ok &= good
    print(f"batch {batch}: |Q13-Q6| {dq:.1e} |R13-R6| {dr:.1e}  vs oracle q {eq:.1e} r {er:.1e} nan {nan} {'OK' if good else 'FAIL'}", flush=True)
    dA.free()
print("ALL OK" if ok else "FAILED")
