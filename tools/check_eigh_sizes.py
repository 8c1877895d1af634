import sys
import numpy as np
sys.path.insert(0, '/root/repo')
from linalg_b200 import _native as nat
import linalg_b200 as lb
ctx = nat.Context(0)
for n in (200, 256, 258, 400, 1000):
    M = np.random.default_rng(n).standard_normal((n + 50, n)); G = M.T @ M
    dG, dl, dV = ctx.upload(G), ctx.alloc(8 * n), ctx.alloc(8 * n * n)
    ctx.call("lq_eigh_dev", dG.ptr, n, dl.ptr, dV.ptr)
    lam = ctx.download(dl, (n,)); V = ctx.download(dV, (n, n))
    ref = np.linalg.eigvalsh(G)[::-1]
    print(n, "lam", np.max(np.abs(lam - ref)) / ref[0], "orth", np.abs(V.T @ V - np.eye(n)).max(), "resid", np.abs(G @ V - V * lam).max() / ref[0], flush=True)
A = np.random.default_rng(1).standard_normal((3000, 500))
U, s, Vt = lb.svd(A, ctx=ctx)
print("svd 3000x500: recon", np.linalg.norm((U * s) @ Vt - A) / np.linalg.norm(A), "s err", np.max(np.abs(s - np.linalg.svd(A, compute_uv=False)) / s))
