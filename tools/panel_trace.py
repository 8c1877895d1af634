#!/usr/bin/env python3
"""clock64() phase stamps of the st.async panel kernel (256 rows per CTA): where a column step spends its cycles."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
ctx = nat.Context(0)
LDA = 8192
names = ["publish+dots+shfl", "syncthreads", "sum+push", "T(prev)", "mbar wait", "recv+scalar+shfl", "update", "(next)"]
for m in (256, 4096):
    A = np.random.default_rng(m).standard_normal((m, 32))
    buf = np.zeros((m, LDA)); buf[:, :32] = A
    dA = ctx.upload(buf); dV = ctx.upload(np.zeros((m, LDA))); dT = ctx.upload(np.zeros((32, 128)))
    dtr = ctx.upload(np.zeros(2 * 32 * 8))
    for _ in range(3):
        ctx.call("lq_debug_panel_trace", dA.ptr, LDA, dV.ptr, LDA, dT.ptr, 128, m, dtr.ptr)
    tr = ctx.download(dtr, (2, 32, 8), dtype=np.int64)
    for wname, t in (("top warp (CTA 0, warp 0)", tr[0]), ("T warp (last CTA, warp 7)", tr[1])):
        d = np.diff(t, axis=1)                      # phase durations inside a step
        nxt = t[1:, 0] - t[:-1, 7]                  # bookkeeping until the next step starts
        step = t[1:, 0] - t[:-1, 0]
        print(f"m={m} {wname}: cycles per column step: mean {step.mean():.0f} (first 8: {step[:8].mean():.0f}, last 8: {step[-8:].mean():.0f}); whole loop {t[-1,7]-t[0,0]}")
        for k in range(7):
            print(f"    {names[k]:20s} mean {d[:, k].mean():7.0f}  min {d[:, k].min():6d}  max {d[:, k].max():6d}")
        print(f"    {'bookkeeping':20s} mean {nxt.mean():7.0f}")
