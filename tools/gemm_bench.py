#!/usr/bin/env python3
"""Correctness + timing of lq_gemm_dev on the shapes the blocked QR / tall-skinny paths use."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
ctx = nat.Context(0)
rng = np.random.default_rng(0)
def run(ta, tb, M, N, K, alpha=1.0, beta=0.0, check=True, reps=5):
    A = rng.standard_normal((K, M) if ta else (M, K)); B = rng.standard_normal((N, K) if tb else (K, N)); Cm = rng.standard_normal((M, N))
    dA, dB, dC = ctx.upload(A), ctx.upload(B), ctx.upload(Cm)
    ctx.call("lq_gemm_dev", ta, tb, M, N, K, C.c_double(alpha), dA.ptr, A.shape[1], dB.ptr, B.shape[1], C.c_double(beta), dC.ptr, N)
    err = None
    if check:
        got = ctx.download(dC, (M, N))
        want = alpha * (A.T if ta else A) @ (B.T if tb else B) + beta * Cm
        err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
    ms = []
    for _ in range(reps):
        ctx.record(0)
        ctx.call("lq_gemm_dev", ta, tb, M, N, K, C.c_double(alpha), dA.ptr, A.shape[1], dB.ptr, B.shape[1], C.c_double(1.0 if beta else 0.0), dC.ptr, N)
        ctx.record(1); ms.append(ctx.elapsed_ms(0, 1))
    t = min(ms)
    print(f"ta={ta} tb={tb} M={M} N={N} K={K} a={alpha} b={beta}: err {err}  {t:.3f} ms  {2.0*M*N*K/t/1e9:.2f} TFLOP/s", flush=True)
    for b in (dA, dB, dC): b.free()
for args in [(0,0,200,136,48,0.5,2.0),(0,1,130,70,35),(1,1,33,17,9),(0,1,256,384,160,-1.0,1.0),(0,0,130,70,32),(0,0,1000,1000,1000),(0,1,777,333,128)]:
    run(*args)
print("--- hot shapes")
run(0,0,4096,4096,4096)
run(1,0,4096,4096,4096)
run(0,1,4096,4096,4096)
run(0,0,8192,8192,128,-1.0,1.0)
run(1,0,128,8192,8192)
run(0,0,8192,96,32,-1.0,1.0)
run(1,0,32,96,8192)
run(1,0,128,128,8192)
run(0,0,1<<20,128,128, check=False)
run(1,0,128,128,1<<20, check=False)
