#!/usr/bin/env python3
"""SYRK Gram kernel (syrk.cu): shapes against NumPy + timing at 2^20 x 128."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
ctx = nat.Context(0)
worst = 0.0
for (m, n) in [(1024, 128), (4096, 128), (5000, 64), (100003, 128), (70001, 30), (2048, 2), (300, 128), (65536, 126), (9999, 127)]:
    A = np.random.default_rng(m + n).standard_normal((m, n))
    dA, dG = ctx.upload(A), ctx.alloc(8 * n * n)
    ctx.call("lq_memset", dG.ptr, 0xFF, 8 * n * n)
    ctx.call("lq_gram_dev", dA.ptr, m, n, dG.ptr)
    G = ctx.download(dG, (n, n))
    Go = A.T @ A
    err = float(np.max(np.abs(G - Go)) / np.max(np.abs(Go)))
    worst = max(worst, err)
    print((m, n), f"rel err {err:.2e}  symmetric {bool(np.array_equal(G, G.T))}", flush=True)
    dA.free(); dG.free()
assert worst < 1e-13
m, n = 1 << 20, 128
A = np.random.default_rng(6).standard_normal((m, n))
dA, dG = ctx.upload(A), ctx.alloc(8 * n * n)
for it in range(4):
    ctx.record(0); ctx.call("lq_gram_dev", dA.ptr, m, n, dG.ptr); ctx.record(1)
    ms = ctx.elapsed_ms(0, 1)
    print(f"gram 2^20 x 128: {ms:.3f} ms  {2.0 * m * n * n / ms / 1e9:.2f} TFLOP/s (GEMM count), {1.0 * m * n * n / ms / 1e9:.2f} (SYRK count)", flush=True)
