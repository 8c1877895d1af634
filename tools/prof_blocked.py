#!/usr/bin/env python3
"""Driver for the ncu launch list of the blocked QR: python tools/prof_blocked.py [n=4096] [reps=1]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = nat.Context(0)
A = np.random.default_rng(5).standard_normal((n, n))
dA = ctx.upload(A); dQ = ctx.alloc(A.nbytes); dR = ctx.alloc(A.nbytes)
for _ in range(reps):
    ctx.record(0)
    ctx.call("lq_householder_qr_dev", dA.ptr, n, n, dQ.ptr, dR.ptr)
    ctx.record(1)
    print("ms", ctx.elapsed_ms(0, 1), flush=True)
