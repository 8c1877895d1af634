"""The rows SURVEY.md section 8(f) marks "next": callers of the hot path, running on the same device kernels.

``project_onto_colspace`` is the reference's ``linalg/projections.py:15-48`` with the switch its own TODO asks for
(``Q, _ = householder_qr(A); return Q @ (Q.T @ b)``, projections.py:23-25); ``pca`` is ``linalg/svd.py:85-123`` on
the A^T A eigen-route of this package instead of LAPACK's SVD.  All matrix products, factorizations and the
eigen-solve run on the GPU through the C ABI; the host only centres the data and assembles the result tuple.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat
from .qr import householder_qr
from .svd import svd


def _matmul(ctx, A: np.ndarray, B: np.ndarray, trans_a: bool = False) -> np.ndarray:
    """op(A) @ B on the device (lq_gemm_dev, FP64 tensor-core GEMM)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    m = A.shape[1] if trans_a else A.shape[0]
    k = A.shape[0] if trans_a else A.shape[1]
    n = B.shape[1]
    if B.shape[0] != k:
        raise ValueError(f"shapes {A.shape} (trans={trans_a}) and {B.shape} not aligned")
    out = np.empty((m, n))
    if m == 0 or n == 0:
        return out
    if k == 0:
        out[:] = 0.0
        return out
    dA, dB, dC = ctx.upload(A), ctx.upload(B), ctx.alloc(out.nbytes)
    ctx.call("lq_gemm_dev", int(trans_a), 0, m, n, k, C.c_double(1.0), dA.ptr, A.shape[1], dB.ptr, n, C.c_double(0.0), dC.ptr, n)
    ctx.download(dC, out.shape, out=out)
    for b in (dA, dB, dC):
        b.free()
    return out


def project_onto_colspace(A, b, *, ctx=None) -> np.ndarray:
    """Orthogonal projection of ``b`` onto the column space of ``A``; shape (m, k) for b of shape (m, k) or (m,)."""
    ctx = ctx if ctx is not None else nat.default_context()
    A = np.asarray(A, dtype=float)
    b = np.asarray(b, dtype=float)
    if b.ndim == 1:
        b = b[:, None]  # projections.py:31-32
    m, n = A.shape
    if m >= n:
        Q, R = householder_qr(A, ctx=ctx)
        d = np.abs(np.diag(R))
        if n == 0 or d.min() > 1e-12 * max(d.max(), 1e-300):
            return _matmul(ctx, Q, _matmul(ctx, Q, b, trans_a=True))  # Q (Q^T b)
    # dependent columns (the reference falls back to the pseudo-inverse, projections.py:35-37):
    # project onto the left singular vectors that carry the range
    U, s, _ = svd(A, ctx=ctx)
    r = int(np.sum(s > 1e-12 * max(s.max() if s.size else 0.0, 1e-300)))
    Ur = np.ascontiguousarray(U[:, :r])
    return _matmul(ctx, Ur, _matmul(ctx, Ur, b, trans_a=True))


def pca(A, k: int, *, ctx=None):
    """PCA with samples in rows (linalg/svd.py:85-123): returns
    ``(pcs, scores, explained_variance, explained_variance_ratio, total_variance, mean_)``."""
    ctx = ctx if ctx is not None else nat.default_context()
    A = np.asarray(A, dtype=float)
    mean_ = A.mean(axis=0, keepdims=True)
    X = A - mean_
    _, S, Vt = svd(X, ctx=ctx)
    pcs = np.ascontiguousarray(Vt[:k].T)
    scores = _matmul(ctx, X, pcs)
    n_samples = A.shape[0]
    explained_variance = (S[:k] ** 2) / (n_samples - 1)
    total_variance = float(np.sum(S ** 2)) / (n_samples - 1)  # = ||X||_F^2 / (n - 1)
    explained_variance_ratio = explained_variance / total_variance
    return pcs, scores, explained_variance, explained_variance_ratio, total_variance, mean_.ravel()
