#!/usr/bin/env python3
"""Shapes beyond the cluster panel's 8192-row reach (generic multi-launch panel) and odd sizes: randomized invariants."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linalg_b200 as lb
from linalg_b200 import _native as nat
ctx = nat.Context(0)
for (m, n) in [(12000, 600), (9001, 333), (10000, 2000)]:
    A = np.random.default_rng(m + n).standard_normal((m, n))
    t0 = time.perf_counter(); Q, R = lb.householder_qr(A, ctx=ctx); t = time.perf_counter() - t0
    X = np.random.default_rng(1).standard_normal((n, 3))
    AX = A @ X
    print((m, n), f"{t*1e3:.1f} ms", "resid", np.linalg.norm(AX - Q @ (R @ X)) / np.linalg.norm(AX), "orth", np.abs(Q.T @ Q - np.eye(n)).max(),
          "lower0", bool(np.all(np.tril(R, -1) == 0)), "sign", bool(abs(R[0, 0] + np.copysign(np.linalg.norm(A[:, 0]), A[0, 0])) < 1e-9 * abs(R[0, 0])), flush=True)
A = np.random.default_rng(7).standard_normal((20000, 500)); B = np.random.default_rng(8).standard_normal((20000, 7))
x = lb.least_squares_householder_qr(A, B, ctx=ctx)
xr = np.linalg.lstsq(A, B, rcond=None)[0]
print("large lstsq rel err", np.abs(x - xr).max() / np.abs(xr).max(), flush=True)
Q, R = lb.qr(A[:, :300], ctx=ctx)
print("large mgs resid", np.linalg.norm(A[:, :300] - Q @ R) / np.linalg.norm(A[:, :300]), "orth", np.abs(Q.T @ Q - np.eye(300)).max(), flush=True)
