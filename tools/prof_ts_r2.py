#!/usr/bin/env python3
"""Driver for the round-2 ncu captures: FP64 probes (DFMA / DMMA peak denominators), SYRK Gram, SVD and TSQR at 2^20 x 128."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
ctx = nat.Context(0)
print("probe dfma TFLOP/s", ctx.probe(0), "dmma", ctx.probe(1), flush=True)
m, n = 1 << 20, 128
A = np.random.default_rng(6).standard_normal((m, n))
dA, dQ, dR, dG = ctx.upload(A), ctx.alloc(A.nbytes), ctx.alloc(8 * n * n), ctx.alloc(8 * n * n)
ds, dVt = ctx.alloc(8 * n), ctx.alloc(8 * n * n)
for _ in range(2):
    ctx.record(0); ctx.call("lq_gram_dev", dA.ptr, m, n, dG.ptr); ctx.record(1); print("gram ms", ctx.elapsed_ms(0, 1), flush=True)
    ctx.record(0); ctx.call("lq_svd_gram_dev", dA.ptr, m, n, C.c_double(1e-12), dQ.ptr, ds.ptr, dVt.ptr, C.byref(C.c_int(0))); ctx.record(1); print("svd ms", ctx.elapsed_ms(0, 1), flush=True)
    ctx.record(0); ctx.call("lq_tsqr_dev", dA.ptr, m, n, dQ.ptr, dR.ptr); ctx.record(1); print("tsqr ms", ctx.elapsed_ms(0, 1), flush=True)
