#!/usr/bin/env python3
"""Generate tests/golden/hotpath_golden.npz by running the REAL reference.

Run in the build container only (the reference tree does not travel to the GPU
box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [/root/reference]

The reference (BrantleighBunting/linalg) ships no known-answer vectors for
qr / householder_qr / least_squares_* / svd, so the
fixtures are outputs of the unmodified reference functions on seeded float64
inputs.  Inputs replay the reference's own seeded test cases where it has them
(tests/test_svd.py:13-16, 38-42, 60-70; tests/test_qr.py:29) plus the
BASELINE.json shape classes at fixture-friendly sizes.
"""
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

from linalg.qr import (  # noqa: E402
    householder_qr,
    least_squares_householder_qr,
    least_squares_qr,
    qr,
)
from linalg.svd import svd  # noqa: E402
from linalg.utils import random_nonsingular_upper  # noqa: E402

out = {}


def put(name, **arrays):
    for k, v in arrays.items():
        out[f"{name}/{k}"] = np.ascontiguousarray(np.asarray(v, dtype=np.float64))


# ---- Householder / MGS QR ------------------------------------------------
qr_cases = {
    "sq32": (32, 32, 2),
    "sq5": (5, 5, 11),
    "tall8x5": (8, 5, 13),
    "sq20": (20, 20, 40),
    "tall50x10": (50, 10, 60),
    "tall100x10": (100, 10, 7),
    "tall64x16": (64, 16, 21),
    "sq64": (64, 64, 22),
    "tall96x40": (96, 40, 23),
    "one1x1": (1, 1, 24),
    "col7x1": (7, 1, 25),
}
for name, (m, n, seed) in qr_cases.items():
    A = np.random.default_rng(seed).standard_normal((m, n))
    Q, R = householder_qr(A)
    put(f"hh/{name}", A=A, Q=Q, R=R)
    Q, R = qr(A)
    put(f"mgs/{name}", A=A, Q=Q, R=R)
    Q, R = qr(A, reorth=True)
    put(f"mgs_reorth/{name}", A=A, Q=Q, R=R)

# Householder edge cases: a zero column (skip branch, qr.py:79-80), negative
# pivots, an already upper-triangular matrix, integer dtype input.
A = np.random.default_rng(31).standard_normal((12, 6))
A[:, 2] = 0.0
Q, R = householder_qr(A)
put("hh/zero_col", A=A, Q=Q, R=R)
A = np.random.default_rng(32).standard_normal((9, 9))
A[3:, 3] = 0.0  # remaining part of column 3 is exactly zero at step 3? (not after updates) -- generic case
Q, R = householder_qr(A)
put("hh/partial_zero", A=A, Q=Q, R=R)
A = np.triu(np.random.default_rng(33).uniform(-5, 5, (10, 10)))
Q, R = householder_qr(A)
put("hh/upper_tri", A=A, Q=Q, R=R)
A = np.zeros((6, 4))
Q, R = householder_qr(A)
put("hh/all_zero", A=A, Q=Q, R=R)
A = np.random.default_rng(34).integers(-9, 10, (16, 8)).astype(np.float64)
Q, R = householder_qr(A)
put("hh/integers", A=A, Q=Q, R=R)

# batched 32x32 (cfg2 shape class): 24 matrices, seed 2 stream
A = np.random.default_rng(2).standard_normal((24, 32, 32))
put(
    "batched32",
    A=A,
    Q_hh=np.stack([np.ascontiguousarray(householder_qr(a)[0]) for a in A]),
    R_hh=np.stack([householder_qr(a)[1] for a in A]),
    Q_mgs=np.stack([qr(a)[0] for a in A]),
    R_mgs=np.stack([qr(a)[1] for a in A]),
)

# ---- least squares -----------------------------------------------------------
for i in range(3):  # tests/test_qr.py:23-47 style, seeded
    U = random_nonsingular_upper(50, seed=100 + i)
    x_true = np.random.default_rng(200 + i).random(50)
    b = U @ x_true
    put(f"ls/upper50_{i}", A=U, b=b, x_hh=least_squares_householder_qr(U, b), x_mgs=least_squares_qr(U, b))
A = np.random.default_rng(3).standard_normal((2, 256, 64))
B = np.random.default_rng(4).standard_normal((2, 256, 16))
put(
    "ls/cfg3",
    A=A,
    B=B,
    X_hh=np.stack([least_squares_householder_qr(A[i], B[i]) for i in range(2)]),
    X_mgs=np.stack([least_squares_qr(A[i], B[i]) for i in range(2)]),
)
A = np.random.default_rng(5).standard_normal((40, 12))
b = np.random.default_rng(6).standard_normal(40)
put("ls/vec40x12", A=A, b=b, x_hh=least_squares_householder_qr(A, b), x_mgs=least_squares_qr(A, b))
B2 = np.random.default_rng(7).standard_normal((40, 3))
put("ls/mat40x12x3", A=A, b=B2, x_hh=least_squares_householder_qr(A, B2), x_mgs=least_squares_qr(A, B2))

# ---- svd -----------------------------------------------------------------------
for m, n in [(8, 5), (20, 20), (50, 10)]:  # tests/test_svd.py:13-16
    A = np.random.default_rng(seed=m + n).normal(size=(m, n))
    U, s, Vt = svd(A)
    put(f"svd/recon_{m}x{n}", A=A, U=U, s=s, Vt=Vt)
for m, n in [(12, 7), (30, 15)]:  # tests/test_svd.py:38-42
    A = np.random.default_rng(seed=4 * m + n).standard_normal(size=(m, n))
    U, s, Vt = svd(A)
    put(f"svd/np_{m}x{n}", A=A, U=U, s=s, Vt=Vt)
for k in (0, 1, 3):  # tests/test_svd.py:60-70 (U's completion columns are random upstream)
    A = np.random.default_rng(123 + k).normal(size=(10, 7))
    if k:
        A[:, -k:] = 0.0
    np.random.seed(999)
    U, s, Vt = svd(A)
    put(f"svd/rankdef_{k}", A=A, U=U, s=s, Vt=Vt)
A = np.random.default_rng(50).standard_normal((5, 9))  # wide -> transpose branch svd.py:37-39
U, s, Vt = svd(A)
put("svd/wide_5x9", A=A, U=U, s=s, Vt=Vt)
A = np.random.default_rng(51).standard_normal((1024, 128))  # cfg5 shape class, reduced rows
U, s, Vt = svd(A)
put("svd/tall_1024x128", A=A, s=s, Vt=Vt)  # U omitted; checked through invariants
Q, R = qr(A[:1024, :64])
put("tsqr/mgs_1024x64", A=A[:1024, :64], R=R)

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hotpath_golden.npz")
np.savez_compressed(path, **out)
print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)/1e6:.2f} MB")
