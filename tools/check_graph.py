#!/usr/bin/env python3
"""CUDA-graph replay of the blocked QR for small single matrices: bitwise equal to the plain multi-stream launches, timing."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
ctx = nat.Context(0)
for (m, n) in ([(int(a), int(a)) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [(256, 256), (300, 300), (512, 512), (1000, 1000), (1024, 256), (1448, 1448)]):
    A = np.random.default_rng(m + n).standard_normal((m, n))
    dA, dQ, dR = ctx.upload(A), ctx.alloc(A.nbytes), ctx.alloc(8 * n * n)
    outs, ms = [], []
    for it in range(6):
        ctx.call("lq_memset", dQ.ptr, 0xFF, A.nbytes)
        ctx.record(0)
        ctx.call("lq_householder_qr_dev", dA.ptr, m, n, dQ.ptr, dR.ptr)
        ctx.record(1)
        ms.append(ctx.elapsed_ms(0, 1))
        outs.append((ctx.download(dQ, (m, n)), ctx.download(dR, (n, n))))
    same = all(np.array_equal(outs[0][0], o[0]) and np.array_equal(outs[0][1], o[1]) for o in outs[1:])
    Q, R = outs[-1]
    resid = np.linalg.norm(A - Q @ R) / np.linalg.norm(A)
    print(f"{m}x{n}: first (plain) {ms[0]:.3f} ms, capture {ms[1]:.3f} ms, replay best {min(ms[2:]):.3f} ms; bitwise equal {same}; resid {resid:.2e}; launches {ctx.launches()}", flush=True)
    for b in (dA, dQ, dR):
        b.free()
