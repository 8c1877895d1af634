"""Constants and host-side helpers shared by the drop-in API (reference: linalg/utils.py)."""
from __future__ import annotations

import numpy as np

EPS: float = 1e-12  # linalg/utils.py:9 -- the absolute threshold compiled into the kernels (kEps)


def as_f64_matrix(A, name: str = "A") -> np.ndarray:
    """C-contiguous float64 copy-or-view of a 2-D array-like.

    Mirrors the reference's ``A.astype(float, copy=True); m, n = A.shape`` (qr.py:28-29, 70-71):
    anything that is not 2-D raises ``ValueError`` just like the tuple unpacking does upstream.
    The caller's array is never written to.
    """
    arr = np.asarray(A)
    if arr.ndim != 2:
        raise ValueError(f"{name} must be 2-D (got shape {arr.shape}); not enough/too many values to unpack")
    return np.ascontiguousarray(arr, dtype=np.float64)


def as_f64_batch(A, name: str = "A") -> np.ndarray:
    arr = np.asarray(A)
    if arr.ndim != 3:
        raise ValueError(f"{name} must be 3-D (batch, rows, cols); got shape {arr.shape}")
    return np.ascontiguousarray(arr, dtype=np.float64)


def shard_bounds(total: int, nranks: int, rank: int, align: int = 1):
    """Contiguous partition of ``total`` units over ``nranks``.

    Rank r gets ``[lo, hi)``; the first ``total % nranks`` ranks get one extra unit (in units of
    ``align``).  Used for both the batch split (no communication) and the row split (TSQR / Gram).
    """
    if nranks < 1 or not (0 <= rank < nranks):
        raise ValueError(f"bad rank {rank} of {nranks}")
    units = total // align
    base, extra = divmod(units, nranks)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    lo *= align
    hi *= align
    if rank == nranks - 1:
        hi = total  # the tail that is not a multiple of `align`
    return lo, hi
