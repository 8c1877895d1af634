#!/usr/bin/env python3
"""Driver for ncu: the rank-128 update GEMM  C -= V W  (NN, M=N=8192, K=128) and the split-K TN product."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
ctx = nat.Context(0)
M = N = 8192; K = 128
rng = np.random.default_rng(0)
A = rng.standard_normal((M, K)); B = rng.standard_normal((K, N)); Cm = rng.standard_normal((M, N))
dA, dB, dC = ctx.upload(A), ctx.upload(B), ctx.upload(Cm)
for _ in range(3):
    ctx.record(0)
    ctx.call("lq_gemm_dev", 0, 0, M, N, K, C.c_double(-1.0), dA.ptr, K, dB.ptr, N, C.c_double(1.0), dC.ptr, N)
    ctx.record(1)
    ms = ctx.elapsed_ms(0, 1)
print(f"NN {M}x{N}x{K}: {ms:.3f} ms {2.0*M*N*K/ms/1e9:.2f} TFLOP/s")
