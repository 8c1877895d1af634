#!/usr/bin/env python3
"""Randomised shape sweep of the drop-in entry points against the oracle (run on a B200 box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linalg_b200 as lb
from oracle import linalg_oracle as orc
rng = np.random.default_rng(12345)
bad = 0
def report(tag, ok, info):
    global bad
    if not ok:
        bad += 1
        print("FAIL", tag, info, flush=True)
shapes = [(1, 1), (2, 1), (5, 3), (33, 32), (40, 40), (64, 64), (65, 64), (100, 100), (129, 128), (130, 33), (257, 129), (300, 300),
          (500, 257), (513, 512), (700, 64), (1025, 96), (2000, 40)]
shapes += [(int(m), int(min(m, n))) for m, n in zip(rng.integers(2, 900, 14), rng.integers(1, 400, 14))]
for (m, n) in shapes:
    A = rng.standard_normal((m, n))
    Q, R = lb.householder_qr(A); Qo, Ro = orc.householder_qr(A)
    report(f"hh {m}x{n}", orc.rel_max_err(Q, Qo) <= 1e-10 and orc.rel_max_err(R, Ro) <= 1e-10 and np.all(np.tril(R, -1) == 0), (orc.rel_max_err(Q, Qo), orc.rel_max_err(R, Ro)))
    Q, R = lb.qr(A); Qo, Ro = orc.mgs_qr(A)
    report(f"mgs {m}x{n}", orc.rel_max_err(Q, Qo) <= 1e-9 and orc.rel_max_err(R, Ro) <= 1e-10, (orc.rel_max_err(Q, Qo), orc.rel_max_err(R, Ro)))
    k = int(rng.integers(1, 5))
    b = rng.standard_normal((m, k)) if k > 1 else rng.standard_normal(m)
    x = lb.least_squares_householder_qr(A, b); xo = orc.lstsq_householder(A, b)
    report(f"lsq_hh {m}x{n}x{k}", x.shape == xo.shape and orc.rel_max_err(x, xo) <= 1e-8, orc.rel_max_err(x, xo))
    x = lb.least_squares_qr(A, b); xo = orc.lstsq_mgs(A, b)
    report(f"lsq_mgs {m}x{n}x{k}", x.shape == xo.shape and orc.rel_max_err(x, xo) <= 1e-8, orc.rel_max_err(x, xo))
    U, s, Vt = lb.svd(A); so = np.linalg.svd(A, compute_uv=False)
    report(f"svd {m}x{n}", np.max(np.abs(s - so)) <= 1e-9 * so[0] and np.abs(U.T @ U - np.eye(U.shape[1])).max() <= 1e-8 and
           np.linalg.norm((U * s) @ Vt - A) <= 1e-10 * np.linalg.norm(A), (np.max(np.abs(s - so)) / so[0], np.abs(U.T @ U - np.eye(U.shape[1])).max()))
    At = A.T.copy()  # wide input: svd transposes (svd.py:37-39)
    U, s, Vt = lb.svd(At)
    report(f"svd wide {n}x{m}", U.shape == (n, n) and Vt.shape == (n, m) and np.linalg.norm((U * s) @ Vt - At) <= 1e-10 * np.linalg.norm(At), U.shape)
    if n <= 128:  # tsqr is the tall-skinny entry (n <= 128)
        Q, R = lb.tsqr(A)
        report(f"tsqr {m}x{n}", np.all(np.diag(R) > 0) and orc.qr_residual(A, Q, R) <= 1e-12 and orc.orth_error(Q) <= 1e-11, (orc.qr_residual(A, Q, R), orc.orth_error(Q)))
for (batch, m, n) in [(7, 32, 32), (100, 16, 16), (33, 48, 20), (5, 100, 64), (64, 64, 64), (9, 200, 31), (3, 256, 64)]:
    A = rng.standard_normal((batch, m, n))
    Q, R = lb.householder_qr_batched(A); Qo, Ro = orc.householder_qr_batched(A)
    report(f"hh batched {batch}x{m}x{n}", orc.rel_max_err(Q, Qo) <= 1e-10 and orc.rel_max_err(R, Ro) <= 1e-10, (orc.rel_max_err(Q, Qo), orc.rel_max_err(R, Ro)))
    Q, R = lb.qr_batched(A); Qo, Ro = orc.mgs_qr_batched(A)
    report(f"mgs batched {batch}x{m}x{n}", orc.rel_max_err(Q, Qo) <= 1e-9 and orc.rel_max_err(R, Ro) <= 1e-10, (orc.rel_max_err(Q, Qo), orc.rel_max_err(R, Ro)))
print("shape sweep:", "ALL OK" if bad == 0 else f"{bad} FAILURES")
# larger, odd shapes: invariants and LAPACK cross-checks only (the oracle is a Python loop)
bad0 = bad
for (m, n) in [(2049, 2049), (3001, 2999), (5000, 1000), (6200, 6150)]:
    A = rng.standard_normal((m, n))
    Q, R = lb.householder_qr(A)
    Rl = np.linalg.qr(A, mode="r")
    X = rng.standard_normal((n, 3))
    resid = np.linalg.norm(A @ X - Q @ (R @ X)) / np.linalg.norm(A @ X)
    orth = np.linalg.norm(Q.T @ (Q @ X) - X) / np.linalg.norm(X)
    dabs = np.max(np.abs(np.abs(np.diag(R)) - np.abs(np.diag(Rl))) / np.abs(np.diag(Rl)))
    report(f"hh large {m}x{n}", resid <= 1e-12 and orth <= 1e-12 and dabs <= 1e-9 and np.all(np.tril(R, -1) == 0), (resid, orth, dabs))
A = rng.standard_normal((4000, 600)); B = rng.standard_normal((4000, 7))
x = lb.least_squares_householder_qr(A, B); xl = np.linalg.lstsq(A, B, rcond=None)[0]
report("lsq_hh large 4000x600x7", orc.rel_max_err(x, xl) <= 1e-9, orc.rel_max_err(x, xl))
print("large shapes:", "ALL OK" if bad == bad0 else f"{bad - bad0} FAILURES")
