#!/usr/bin/env python3
"""Blocked Householder QR: parity at moderate sizes + timing at 256 .. 8192 (device resident)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
from oracle import linalg_oracle as orc
ctx = nat.Context(0)
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [256, 1024, 4096, 8192]
for (m, n) in [(300, 300), (700, 130), (1000, 1000)]:
    A = np.random.default_rng(m + n).standard_normal((m, n))
    dA, dQ, dR = ctx.upload(A), ctx.alloc(A.nbytes), ctx.alloc(8 * n * n)
    ctx.call("lq_householder_qr_dev", dA.ptr, m, n, dQ.ptr, dR.ptr)
    Q, R = ctx.download(dQ, (m, n)), ctx.download(dR, (n, n))
    Qo, Ro = orc.householder_qr(A)
    print((m, n), "q", orc.rel_max_err(Q, Qo), "r", orc.rel_max_err(R, Ro), "resid", orc.qr_residual(A, Q, R), "orth", orc.orth_error(Q), flush=True)
for n in sizes:
    A = np.random.default_rng(5).standard_normal((n, n))
    dA, dQ, dR = ctx.upload(A), ctx.alloc(A.nbytes), ctx.alloc(A.nbytes)
    ms = []
    l0 = ctx.launches()
    for _ in range(4):
        ctx.record(0); ctx.call("lq_householder_qr_dev", dA.ptr, n, n, dQ.ptr, dR.ptr); ctx.record(1)
        ms.append(ctx.elapsed_ms(0, 1))
    t = min(ms[1:])
    print(f"{n}x{n}: {t:.2f} ms  F_QR {8/3*n**3/t/1e9:.2f} TFLOP/s  launches/call {(ctx.launches()-l0)//4}", flush=True)
    if n >= 4096:
        Q = ctx.download(dQ, (n, n)); R = ctx.download(dR, (n, n))
        X = np.random.default_rng(6).standard_normal((n, 3))
        print("   probe resid", np.linalg.norm(A @ X - Q @ (R @ X)) / np.linalg.norm(A @ X), "orth", np.linalg.norm(Q.T @ (Q @ X) - X) / np.linalg.norm(X), flush=True)
    for b in (dA, dQ, dR): b.free()
