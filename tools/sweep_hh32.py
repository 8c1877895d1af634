#!/usr/bin/env python3
"""Time every hh32 kernel variant (and mgs32) on 2^18 matrices and check them against the oracle."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
from oracle import linalg_oracle as orc

variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else list(range(1, 13))
batch = 1 << 18
ctx = nat.Context(0)
A = np.random.default_rng(2).standard_normal((1 << 14, 32, 32))
per = A.nbytes
reps = batch // A.shape[0]
dA, dQ, dR = ctx.alloc(per * reps), ctx.alloc(per * reps), ctx.alloc(per * reps)
ctx.call("lq_memcpy_h2d", dA.ptr, A.ctypes.data, per)
for r in range(1, reps):
    ctx.call("lq_memcpy_d2d", dA.ptr + r * per, dA.ptr, per)
ctx.sync()
Qo, Ro = orc.householder_qr_batched(A[:64])
out = {}
print("dfma latency cycles", ctx.probe(4))
for v in variants:
    try:
        for _ in range(3):
            ctx.call("lq_householder_qr_batched_dev", dA.ptr, batch, 32, 32, dQ.ptr, dR.ptr, v)
        ctx.sync()
        ms = []
        for _ in range(5):
            ctx.record(0)
            ctx.call("lq_householder_qr_batched_dev", dA.ptr, batch, 32, 32, dQ.ptr, dR.ptr, v)
            ctx.record(1)
            ms.append(ctx.elapsed_ms(0, 1))
        Q = np.empty((64, 32, 32)); R = np.empty((64, 32, 32))
        ctx.call("lq_memcpy_d2h", Q.ctypes.data, dQ.ptr + (reps - 1) * per, Q.nbytes)
        ctx.call("lq_memcpy_d2h", R.ctypes.data, dR.ptr + (reps - 1) * per, R.nbytes)
        ctx.sync()
        t = min(ms)
        out[v] = dict(ms=t, mps=batch / t * 1e3, gbs=batch * 24576 / t / 1e6, q=orc.rel_max_err(Q, Qo), r=orc.rel_max_err(R, Ro))
        print("variant", v, {k: (round(x, 3) if x > 1e-3 else x) for k, x in out[v].items()}, flush=True)
    except Exception as e:
        print("variant", v, "FAILED", e, flush=True)
dI = ctx.alloc(4 * batch)
Qm, Rm = orc.mgs_qr_batched(A[:64])
for reorth in (0, 1):
    ms = []
    for _ in range(6):
        ctx.record(0)
        ctx.call("lq_mgs_qr_batched_dev", dA.ptr, batch, 32, 32, reorth, dQ.ptr, dR.ptr, dI.ptr)
        ctx.record(1)
        ms.append(ctx.elapsed_ms(0, 1))
    Q = np.empty((64, 32, 32)); R = np.empty((64, 32, 32))
    ctx.call("lq_memcpy_d2h", Q.ctypes.data, dQ.ptr, Q.nbytes)
    ctx.call("lq_memcpy_d2h", R.ctypes.data, dR.ptr, R.nbytes)
    ctx.sync()
    t = min(ms[1:])
    extra = dict(q=orc.rel_max_err(Q, Qm), r=orc.rel_max_err(R, Rm)) if reorth == 0 else {}
    print("mgs reorth", reorth, dict(ms=round(t, 3), mps=round(batch / t * 1e3), gbs=round(batch * 24576 / t / 1e6), **extra), flush=True)
