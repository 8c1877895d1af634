#!/usr/bin/env python3
"""The reference's QR benchmark table (linalg/benchmark_qr.py:14-68 methodology: seed 0, randn inputs, shapes
(300,300), (1000,1000), (5000,1000), min wall time of 5 calls, residual relative to np.linalg.lstsq, orth_err =
||Q^T Q - I||_inf) for the drop-in entry points of linalg_b200 -- NumPy in, NumPy out, PCIe copies included.

    python tools/benchmark_qr_table.py            # on a B200 box
"""
import os, sys, time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linalg_b200 as lb

np.random.seed(0)
REPEATS = 5
sizes = [(300, 300), (1000, 1000), (5000, 1000)]


def wall(f, *a, **k):
    t0 = time.perf_counter(); f(*a, **k); return time.perf_counter() - t0


rows = []
lb.householder_qr(np.random.randn(64, 64)); lb.qr(np.random.randn(64, 64))  # context + kernels warm
for m, n in sizes:
    A = np.random.randn(m, n); b = np.random.randn(m)
    t_np = min(wall(np.linalg.lstsq, A, b, rcond=None) for _ in range(REPEATS))
    x_ref, *_ = np.linalg.lstsq(A, b, rcond=None)
    r_ref = np.linalg.norm(A @ x_ref - b, np.inf)
    for name, fqr, fls in (("MGS-QR", lb.qr, lb.least_squares_qr), ("HH-QR", lb.householder_qr, lb.least_squares_householder_qr)):
        t = min(wall(fqr, A) for _ in range(REPEATS))
        Q, R = fqr(A)
        ortho = np.linalg.norm(Q.T @ Q - np.eye(n), np.inf)
        x = fls(A, b)
        r = np.linalg.norm(A @ x - b, np.inf)
        rows.append((name, f"{m}x{n}", t, t / t_np, r / max(r_ref, 1e-300), ortho))
print("| kernel | size | sec | sec/NumPy lstsq | residual/NumPy | orth_err |\n|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r[0]} | {r[1]} | {r[2]:.5f} | {r[3]:.3f} | {r[4]:.3f} | {r[5]:.2e} |")
