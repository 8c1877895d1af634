#!/usr/bin/env python3
"""Driver for the ncu launch list of the tall-skinny paths: one lq_tsqr_dev and one lq_svd_gram_dev at 2^20 x 128."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
ctx = nat.Context(0)
m, n = 1 << 20, 128
A = np.random.default_rng(6).standard_normal((m, n))
dA, dQ, dR = ctx.upload(A), ctx.alloc(A.nbytes), ctx.alloc(8 * n * n)
ds, dVt = ctx.alloc(8 * n), ctx.alloc(8 * n * n)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(reps):
    ctx.record(0); ctx.call("lq_tsqr_dev", dA.ptr, m, n, dQ.ptr, dR.ptr); ctx.record(1); print("tsqr ms", ctx.elapsed_ms(0, 1))
    ctx.record(0); ctx.call("lq_svd_gram_dev", dA.ptr, m, n, C.c_double(1e-12), dQ.ptr, ds.ptr, dVt.ptr, C.byref(C.c_int(0))); ctx.record(1); print("svd ms", ctx.elapsed_ms(0, 1))
