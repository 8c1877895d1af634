"""CPU oracle for the dense-factorisation hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the five reference routines on the hot
path (reference paths are relative to the upstream repo root):

    mgs_qr                 <- linalg/qr.py:14-49    (``qr``)
    householder_qr         <- linalg/qr.py:52-100   (``householder_qr``)
    lstsq_mgs              <- linalg/qr.py:103-119  (``least_squares_qr``)
    lstsq_householder      <- linalg/qr.py:122-134  (``least_squares_householder_qr``)
    svd_gram               <- linalg/svd.py:10-82   (``svd``)
    EPS                    <- linalg/utils.py:9

It issues the same NumPy/BLAS calls in the same order as the reference, so on
identical float64 input the results are bit-identical to the reference's
(``tests/golden/make_golden.py`` generates fixtures by importing the real
reference, and ``tests/test_oracle_golden.py`` pins this file against them
with ``np.array_equal``).  Parity status: PINNED (against reference outputs
produced in the build container; the reference ships no golden vectors of its
own for this path).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  Nothing in
``linalg_b200/`` imports it; the product has no CPU compute path.
"""

from __future__ import annotations

import numpy as np

EPS: float = 1e-12  # linalg/utils.py:9 -- absolute threshold in both QR routines


# --------------------------------------------------------------------------
# a2: modified Gram-Schmidt                              linalg/qr.py:14-49
# --------------------------------------------------------------------------
def mgs_qr(A, reorth: bool = False):
    """Left-looking modified Gram-Schmidt, ``diag(R) > 0``.

    Quirks kept on purpose: with ``reorth=True`` the second
    sweep runs over the Q of the first sweep *and overwrites R*, so the R that
    comes back belongs to the second sweep (close to I) and ``Q @ R != A``.
    Raises ``ValueError("Input vectors are linearly dependent")`` when a
    column's remaining norm drops below ``EPS`` (qr.py:40-41).
    """
    work = np.asarray(A).astype(float, copy=True)  # qr.py:28
    rows, cols = work.shape  # ndim != 2 -> ValueError, as upstream (qr.py:29)
    Q = np.zeros_like(work)
    R = np.zeros((cols, cols))

    def sweep(src):  # qr.py:33-43
        for j in range(cols):
            v = src[:, j].copy()
            for k in range(j):
                R[k, j] = Q[:, k] @ v
                v -= R[k, j] * Q[:, k]
            R[j, j] = np.linalg.norm(v)
            if R[j, j] < EPS:
                raise ValueError("Input vectors are linearly dependent")
            Q[:, j] = v / R[j, j]
        return Q.copy()

    Q = sweep(work)  # qr.py:45
    if reorth:
        Q = sweep(Q)  # qr.py:46-47
    return Q, R


# --------------------------------------------------------------------------
# a1: Householder QR                                     linalg/qr.py:52-100
# --------------------------------------------------------------------------
def householder_qr(A):
    """Thin QR by ``n`` unit-norm reflectors ``H = I - 2 w w^T``.

    Sign convention: ``R[j, j] = -copysign(||x||, x[0])`` for
    every column including the last one of a square matrix; a column whose
    remaining norm is below ``EPS`` (absolute) is skipped.  Q is accumulated in
    a full ``max(m, n)`` square and sliced, R's strict lower triangle is set to
    exact zeros.  ``m < n`` ends in a matmul shape ``ValueError`` like upstream.
    """
    work = np.asarray(A).astype(float, copy=True)  # qr.py:70
    rows, cols = work.shape
    Qfull = np.eye(max(rows, cols))  # qr.py:72
    R = work.copy().astype(float)  # qr.py:73

    for j in range(cols):  # qr.py:75
        x = R[j:, j]
        nx = np.linalg.norm(x)
        if nx < EPS:  # qr.py:79-80
            continue
        w = x.copy()
        w[0] += np.copysign(nx, x[0])  # qr.py:83
        w /= np.linalg.norm(w)
        w = w.reshape(-1, 1)
        two = 2
        R[j:, :] -= two * w @ (w.T @ R[j:, :])  # qr.py:89
        Qfull[:, j:] -= Qfull[:, j:] @ w @ (two * w).T  # qr.py:91

    Q = Qfull[:, :cols]  # qr.py:94
    R[np.tril_indices(cols, -1)] = 0.0  # qr.py:97
    R = R[:cols, :cols]  # qr.py:99
    return Q, R


# --------------------------------------------------------------------------
# a4 / a3: least squares                                linalg/qr.py:103-134
# --------------------------------------------------------------------------
def lstsq_mgs(A, b):
    """``x = R^-1 (Q^T b)`` with MGS factors; result is always ravel()ed (qr.py:119)."""
    rows, cols = A.shape
    Q, R = mgs_qr(A)
    y = Q.T @ b
    if rows > cols:
        x = np.linalg.solve(R[:cols, :], y[:cols])  # qr.py:116
    else:
        x = np.linalg.solve(R, y)  # qr.py:118
    return x.ravel()


def lstsq_householder(A, b):
    """``x = R^-1 (Q^T b)`` with Householder factors; shape ``(n,)`` or ``(n, k)``."""
    Q, R = householder_qr(A)
    y = Q.T @ b  # qr.py:133
    return np.linalg.solve(R, y)  # qr.py:134


# --------------------------------------------------------------------------
# a5: economy SVD through the Gram matrix                linalg/svd.py:10-82
# --------------------------------------------------------------------------
def svd_gram(A, tol: float = 1e-12, rng=None):
    """Economy SVD via ``eigh(A^T A)``; returns ``(U, s, Vt)``.

    ``rng``: the reference draws the rank-deficient completion from the global
    ``np.random`` state (svd.py:69); pass a ``np.random.RandomState`` to make
    the oracle reproducible, default ``None`` keeps the upstream behaviour.
    """
    M = np.asarray(A, dtype=float)
    rows, cols = M.shape
    if rows < cols:  # svd.py:37-39
        Vt, s, Ut = svd_gram(M.T, tol, rng)
        return Ut.T, s, Vt.T

    gram = M.T @ M  # svd.py:42
    lam, V = np.linalg.eigh(gram)  # svd.py:46
    order = np.argsort(lam)[::-1]  # svd.py:49
    lam = lam[order]
    V = V[:, order]
    s = np.sqrt(np.clip(lam, 0.0, None))  # svd.py:54
    rank = np.sum(s > tol)  # svd.py:57

    left = []
    for j, sigma in enumerate(s):  # svd.py:61-64
        if sigma > tol:
            left.append(M @ V[:, j] / sigma)

    if rank < cols:  # svd.py:67-76
        draw = np.random.randn if rng is None else rng.randn
        C, _ = np.linalg.qr(draw(rows, cols - rank))
        for u in left:
            C -= u[:, None] * (u @ C)
        C, _ = np.linalg.qr(C)
        left.extend(C[:, k] for k in range(cols - rank))

    return np.column_stack(left), s, V.T  # svd.py:79-82


# --------------------------------------------------------------------------
# Batched / tall-skinny definitions (new surface):
# "the reference function applied independently to each A[i]".
# --------------------------------------------------------------------------
def householder_qr_batched(A):
    A = np.asarray(A, dtype=float)
    out = [householder_qr(A[i]) for i in range(A.shape[0])]
    Q = np.stack([np.ascontiguousarray(q) for q, _ in out])
    R = np.stack([r for _, r in out])
    return Q, R


def mgs_qr_batched(A, reorth: bool = False):
    A = np.asarray(A, dtype=float)
    out = [mgs_qr(A[i], reorth) for i in range(A.shape[0])]
    return np.stack([q for q, _ in out]), np.stack([r for _, r in out])


def lstsq_householder_batched(A, B):
    return np.stack([lstsq_householder(A[i], B[i]) for i in range(A.shape[0])])


def lstsq_mgs_batched(A, B):
    return np.stack([lstsq_mgs(A[i], B[i]) for i in range(A.shape[0])])


def tsqr_reference(A):
    """a7: thin QR with ``diag(R) > 0`` -- the MGS convention (qr.py:39-42).

    At sizes where ``mgs_qr`` is affordable the tests use it directly; this
    helper gives the same factorisation (up to rounding) from LAPACK for the
    larger row counts by flipping signs so that the diagonal is positive.
    """
    Q, R = np.linalg.qr(np.asarray(A, dtype=float))
    sgn = np.where(np.diag(R) < 0, -1.0, 1.0)
    return Q * sgn, sgn[:, None] * R


# --------------------------------------------------------------------------
# Tolerance helpers implementing BASELINE.json's north_star bars.
# --------------------------------------------------------------------------
def rel_max_err(got, want):
    """max |got - want| / max |want|  (the "relative 1e-10" bar for Q, R, x, s)."""
    want = np.asarray(want)
    scale = np.max(np.abs(want)) if want.size else 1.0
    scale = scale if scale > 0 else 1.0
    return float(np.max(np.abs(np.asarray(got) - want)) / scale) if want.size else 0.0


def qr_residual(A, Q, R):
    """||A - QR||_F / ||A||_F   (bar: <= 1e-12)."""
    A = np.asarray(A, dtype=float)
    den = np.linalg.norm(A)
    return float(np.linalg.norm(A - Q @ R) / (den if den > 0 else 1.0))


def orth_error(Q):
    """max |Q^T Q - I|   (bar: <= 1e-12)."""
    n = Q.shape[-1]
    G = np.swapaxes(Q, -1, -2) @ Q
    return float(np.max(np.abs(G - np.eye(n))))
