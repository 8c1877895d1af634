#!/usr/bin/env python3
"""Tall-skinny paths: parity + timing of tsqr / svd_gram (device resident)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linalg_b200 as lb
from linalg_b200 import _native as nat
from oracle import linalg_oracle as orc
ctx = nat.Context(0)
for (m, n) in [(1000, 17), (65536, 128), (300001, 64)]:
    A = np.random.default_rng(m % 1000 + n).standard_normal((m, n))
    Q, R = lb.tsqr(A, ctx=ctx)
    Qo, Ro = orc.tsqr_reference(A)
    print((m, n), "r", orc.rel_max_err(R, Ro), "q", orc.rel_max_err(Q, Qo), "resid", orc.qr_residual(A, Q, R), "orth", orc.orth_error(Q), flush=True)
# ill-conditioned: columns scaled over 12 decades -> must fall back to the Householder tree and stay accurate
A = np.random.default_rng(3).standard_normal((20000, 32)) * np.logspace(0, -9, 32)
A[:, 5] = A[:, 4] * (1 + 1e-9) + 1e-9 * A[:, 6]
Q, R = lb.tsqr(A, ctx=ctx)
print("ill-cond: resid", orc.qr_residual(A, Q, R), "orth", orc.orth_error(Q), "diag>0", bool(np.all(np.diag(R) > 0)), flush=True)
M0 = np.random.default_rng(9).standard_normal((4096, 128)); G0 = M0.T @ M0
dG, dl, dV = ctx.upload(G0), ctx.alloc(8 * 128), ctx.alloc(8 * 128 * 128)
ms = []
for _ in range(4):
    ctx.record(0); ctx.call("lq_eigh_dev", dG.ptr, 128, dl.ptr, dV.ptr); ctx.record(1); ms.append(ctx.elapsed_ms(0, 1))
lam = ctx.download(dl, (128,)); ref = np.linalg.eigvalsh(G0)[::-1]
print(f"eigh 128: {min(ms[1:]):.2f} ms, max rel err {np.max(np.abs(lam-ref)/ref):.2e}", flush=True)
m, n = 1 << 20, 128
A = np.random.default_rng(6).standard_normal((m, n))
dA, dQ, dR = ctx.upload(A), ctx.alloc(A.nbytes), ctx.alloc(8 * n * n)
ds, dVt = ctx.alloc(8 * n), ctx.alloc(8 * n * n)
for name, fn in (("tsqr", lambda: ctx.call("lq_tsqr_dev", dA.ptr, m, n, dQ.ptr, dR.ptr)),
                 ("svd_gram", lambda: ctx.call("lq_svd_gram_dev", dA.ptr, m, n, C.c_double(1e-12), dQ.ptr, ds.ptr, dVt.ptr, C.byref(C.c_int(0))))):
    ms = []
    for _ in range(5):
        ctx.record(0); fn(); ctx.record(1); ms.append(ctx.elapsed_ms(0, 1))
    print(name, f"2^20 x 128: {min(ms[1:]):.2f} ms", flush=True)
