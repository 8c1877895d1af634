#!/usr/bin/env python3
"""Driver for ncu: batched Householder least squares, cfg3 shape (256x64, 16 rhs)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
nsys = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = nat.Context(0)
A = np.random.default_rng(3).standard_normal((nsys, 256, 64)); B = np.random.default_rng(4).standard_normal((nsys, 256, 16))
dA, dB, dX = ctx.upload(A), ctx.upload(B), ctx.alloc(8 * nsys * 64 * 16)
for _ in range(3):
    ctx.record(0); ctx.call("lq_lstsq_householder_batched_dev", dA.ptr, dB.ptr, nsys, 256, 64, 16, dX.ptr); ctx.record(1)
    ms = ctx.elapsed_ms(0, 1)
print(f"{nsys} systems: {ms:.3f} ms  {nsys/ms/1e3:.3f} M sys/s  {nsys*2905429/ms/1e9:.2f} TFLOP/s")
