#!/usr/bin/env python3
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linalg_b200 import _native as nat
ctx = nat.Context(0)
M0 = np.random.default_rng(9).standard_normal((4096, 128)); G0 = M0.T @ M0
dG, dl, dV = ctx.upload(G0), ctx.alloc(8 * 128), ctx.alloc(8 * 128 * 128)
for _ in range(2):
    ctx.call("lq_eigh_dev", dG.ptr, 128, dl.ptr, dV.ptr)
ctx.sync()
