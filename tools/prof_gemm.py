#!/usr/bin/env python3
"""Driver for ncu: the two GEMM shapes of a 128-wide block-reflector application at full size,
W = V^T C (TN, 128 x 8192 x 8192, split-K) and C -= V W (NN, 8192 x 8192 x 128), plus a square 4096^3 product."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
ctx = nat.Context(0)
rng = np.random.default_rng(0)
def run(ta, tb, M, N, K, alpha, beta, reps=2):
    A = rng.standard_normal((K, M) if ta else (M, K)); B = rng.standard_normal((N, K) if tb else (K, N)); Cm = rng.standard_normal((M, N))
    dA, dB, dC = ctx.upload(A), ctx.upload(B), ctx.upload(Cm)
    for _ in range(reps):
        ctx.record(0)
        ctx.call("lq_gemm_dev", ta, tb, M, N, K, C.c_double(alpha), dA.ptr, A.shape[1], dB.ptr, B.shape[1], C.c_double(beta), dC.ptr, N)
        ctx.record(1); ms = ctx.elapsed_ms(0, 1)
    print(f"ta={ta} tb={tb} {M}x{N}x{K}: {ms:.3f} ms {2.0*M*N*K/ms/1e9:.2f} TFLOP/s", flush=True)
    for b in (dA, dB, dC): b.free()
run(0, 0, 8192, 8192, 128, -1.0, 1.0)
run(1, 0, 128, 8192, 8192, 1.0, 0.0)
run(0, 0, 4096, 4096, 4096, 1.0, 0.0)
