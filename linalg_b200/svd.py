"""Drop-in replacement for ``linalg/svd.py::svd`` -- economy SVD through the A^T A eigen-route.

Gram matrix, Jacobi eigen-solver, descending sort, ``s = sqrt(max(lambda, 0))``, ``U = A V / s``
all run on the device (linalg/svd.py:42-64).  Eigenvector signs are whatever the eigen-solver
produces (the reference inherits LAPACK's), so U and V agree with the reference up to a per-column
sign, which is also how the reference's own test compares them (tests/test_svd.py:31-35).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat


def svd(A, tol: float = 1e-12, *, ctx=None, seed=None):
    """Returns ``(U, s, Vt)``: U (m, n) orthonormal columns, s (n,) descending, Vt (n, n), for m >= n;
    a wide matrix is handled by transposing and swapping roles (linalg/svd.py:37-39).

    ``seed`` only matters for rank-deficient input (linalg/svd.py:67-76): ``None`` (default) draws the candidate
    directions of the orthonormal completion from the global ``np.random`` exactly like upstream (not reproducible);
    an integer generates them on the device (Philox4x32-10), deterministically and without the m x (n - rank) upload."""
    A = np.asarray(A, dtype=float)
    if A.ndim != 2:
        raise ValueError(f"A must be 2-D (got shape {A.shape}); not enough/too many values to unpack")
    m, n = A.shape
    if m < n:
        Vt, s, Ut = svd(A.T, tol, ctx=ctx, seed=seed)
        return Ut.T, s, Vt.T
    A = np.ascontiguousarray(A, dtype=np.float64)
    ctx = ctx if ctx is not None else nat.default_context()
    U = np.empty((m, n))
    s = np.empty(n)
    Vt = np.empty((n, n))
    if n == 0:
        return U, s, Vt
    rank = C.c_int(0)
    ctx.call("lq_svd_gram", A.ctypes.data, m, n, float(tol), U.ctypes.data, s.ctypes.data, Vt.ctypes.data, C.byref(rank))
    r = int(rank.value)
    if r < n:
        # linalg/svd.py:67-76 -- complete U with an orthonormal basis of the complement.  The
        # candidate directions are drawn on the host exactly like upstream (global np.random);
        # the projection and both orthonormalisations run on the device (lq_svd_complete).
        if seed is None:
            Z = np.ascontiguousarray(np.random.randn(m, n - r))
            ctx.call("lq_svd_complete", U.ctypes.data, m, n, r, Z.ctypes.data)
        else:
            ctx.call("lq_svd_complete_seeded", U.ctypes.data, m, n, r, int(seed) & 0xFFFFFFFFFFFFFFFF)
    return U, s, Vt
