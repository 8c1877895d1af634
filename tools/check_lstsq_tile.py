#!/usr/bin/env python3
"""Round-2 K3 kernel (lstsq_tile.cuh): shape sweep against LAPACK + cfg3 timing, device-resident.

    python tools/check_lstsq_tile.py [nsys]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat  # noqa: E402

ctx = nat.Context(0)
rng = np.random.default_rng(11)
worst = 0.0
for (b, m, n, k) in [(5, 256, 64, 16), (3, 64, 64, 16), (4, 33, 17, 3), (7, 50, 50, 1), (2, 100, 10, 16), (3, 300, 64, 9),
                     (2, 257, 63, 15), (9, 8, 8, 8), (3, 40, 1, 1), (2, 1000, 48, 16), (65, 256, 64, 16)]:
    A = rng.standard_normal((b, m, n))
    B = rng.standard_normal((b, m, k))
    dA, dB, dX, dI = ctx.upload(A), ctx.upload(B), ctx.alloc(8 * b * n * k), ctx.alloc(4 * b)
    for name, extra in (("lq_lstsq_householder_batched_dev", ()), ("lq_lstsq_mgs_batched_dev", (dI.ptr,))):
        ctx.call("lq_memset", dX.ptr, 0xFF, 8 * b * n * k)
        ctx.call(name, dA.ptr, dB.ptr, b, m, n, k, dX.ptr, *extra)
        X = ctx.download(dX, (b, n, k))
        err = 0.0
        for i in range(b):
            Xo = np.linalg.lstsq(A[i], B[i], rcond=None)[0]
            err = max(err, float(np.max(np.abs(X[i] - Xo)) / np.max(np.abs(Xo))))
        worst = max(worst, err)
        print(f"{name[3:22]:20s} b={b:3d} {m:4d}x{n:2d} k={k:2d}  rel err vs LAPACK {err:.2e}", flush=True)
    for x in (dA, dB, dX, dI):
        x.free()
# randomised shapes (every m mod 32, n mod 8, nrhs): ragged last blocks, partial panels, partial right-hand-side tiles
frng = np.random.default_rng(2026)
for trial in range(60):
    n = int(frng.integers(1, 65)); k = int(frng.integers(1, 17)); m = int(frng.integers(n, 700)); b = int(frng.integers(1, 40))
    A = frng.standard_normal((b, m, n)) * 10.0 ** frng.uniform(-3, 3)
    B = frng.standard_normal((b, m, k))
    dA, dB, dX, dI = ctx.upload(A), ctx.upload(B), ctx.alloc(8 * b * n * k), ctx.alloc(4 * b)
    ctx.call("lq_memset", dX.ptr, 0xFF, 8 * b * n * k)
    ctx.call("lq_lstsq_householder_batched_info_dev", dA.ptr, dB.ptr, b, m, n, k, dX.ptr, dI.ptr)
    X = ctx.download(dX, (b, n, k))
    assert not ctx.download(dI, (b,), dtype=np.int32).any()
    err = 0.0
    for i in range(b):
        Xo = np.linalg.lstsq(A[i], B[i], rcond=None)[0]
        err = max(err, float(np.max(np.abs(X[i] - Xo)) / np.max(np.abs(Xo))))
    cond = max(np.linalg.cond(A[i]) for i in range(min(b, 4)))
    assert err < 1e-10 * max(1.0, cond / 1e3), (trial, b, m, n, k, err, cond)
    worst = max(worst, err / max(1.0, cond / 1e3))
    for x in (dA, dB, dX, dI):
        x.free()
print("60 random shapes ok; worst scaled rel err", worst, flush=True)
# dependent columns -> info (MGS semantics), zero column -> exactly singular
A = rng.standard_normal((4, 256, 64))
A[1, :, 40] = A[1, :, 3] * 2.0
A[3, :, 7] = 0.0
B = rng.standard_normal((4, 256, 16))
dA, dB, dX, dI = ctx.upload(A), ctx.upload(B), ctx.alloc(8 * 4 * 64 * 16), ctx.alloc(16)
ctx.call("lq_lstsq_mgs_batched_dev", dA.ptr, dB.ptr, 4, 256, 64, 16, dX.ptr, dI.ptr)
print("info (expect [0, 41, 0, 8]):", ctx.download(dI, (4,), dtype=np.int32))
print("worst rel err", worst)
assert worst < 1e-10

nsys = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
uniq = 2048
A = np.random.default_rng(3).standard_normal((uniq, 256, 64))
B = np.random.default_rng(4).standard_normal((uniq, 256, 16))
dA, dB = ctx.alloc(8 * nsys * 256 * 64), ctx.alloc(8 * nsys * 256 * 16)
for r in range(nsys // uniq):
    ctx.call("lq_memcpy_h2d", dA.ptr + r * A.nbytes, A.ctypes.data, A.nbytes)
    ctx.call("lq_memcpy_h2d", dB.ptr + r * B.nbytes, B.ctypes.data, B.nbytes)
ctx.sync()
dX = ctx.alloc(8 * nsys * 64 * 16)
for it in range(4):
    ctx.record(0)
    ctx.call("lq_lstsq_householder_batched_dev", dA.ptr, dB.ptr, nsys, 256, 64, 16, dX.ptr)
    ctx.record(1)
    ms = ctx.elapsed_ms(0, 1)
    print(f"{nsys} systems: {ms:.3f} ms  {nsys / ms / 1e3:.3f} M sys/s  {nsys * 2905429 / ms / 1e9:.2f} TFLOP/s", flush=True)
