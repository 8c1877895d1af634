"""The oracle (oracle/linalg_oracle.py) pinned against outputs of the REAL reference.

``tests/golden/hotpath_golden.npz`` was produced by ``tests/golden/make_golden.py`` importing the
unmodified reference (linalg/qr.py, linalg/svd.py) in the build container.  The reference has no
golden vectors of its own on this path, so these fixtures are the pin.
The oracle issues the same NumPy calls in the same order, hence the comparison is tight
(1e-13 relative; the BLAS in the box may differ from the one that generated the fixtures by
summation order only).
"""
import numpy as np
import pytest

from conftest import golden_cases
from oracle import linalg_oracle as orc

TIGHT = 2e-13


def _close(got, want, tol=TIGHT):
    assert got.shape == want.shape
    assert orc.rel_max_err(got, want) <= tol


@pytest.mark.parametrize("name", golden_cases("hh"))
def test_householder_matches_reference(golden, name):
    A = golden[f"hh/{name}/A"]
    Q, R = orc.householder_qr(A)
    _close(np.ascontiguousarray(Q), golden[f"hh/{name}/Q"])
    _close(R, golden[f"hh/{name}/R"])
    assert np.all(np.tril(R, -1) == 0.0)  # qr.py:97


@pytest.mark.parametrize("name", golden_cases("mgs"))
def test_mgs_matches_reference(golden, name):
    A = golden[f"mgs/{name}/A"]
    Q, R = orc.mgs_qr(A)
    _close(Q, golden[f"mgs/{name}/Q"])
    _close(R, golden[f"mgs/{name}/R"])
    assert np.all(np.diag(R) > 0)  # qr.py:39


@pytest.mark.parametrize("name", golden_cases("mgs_reorth"))
def test_mgs_reorth_quirk_matches_reference(golden, name):
    A = golden[f"mgs_reorth/{name}/A"]
    Q, R = orc.mgs_qr(A, reorth=True)
    _close(Q, golden[f"mgs_reorth/{name}/Q"])
    _close(R, golden[f"mgs_reorth/{name}/R"], 1e-12)  # second-sweep R ~ I (qr.py:46-47)


def test_batched32_matches_reference(golden):
    A = golden["batched32/A"]
    Q, R = orc.householder_qr_batched(A)
    _close(Q, golden["batched32/Q_hh"])
    _close(R, golden["batched32/R_hh"])
    Q, R = orc.mgs_qr_batched(A)
    _close(Q, golden["batched32/Q_mgs"])
    _close(R, golden["batched32/R_mgs"])


@pytest.mark.parametrize("name", golden_cases("ls"))
def test_least_squares_matches_reference(golden, name):
    A = golden[f"ls/{name}/A"]
    if name == "cfg3":
        B = golden["ls/cfg3/B"]
        _close(orc.lstsq_householder_batched(A, B), golden["ls/cfg3/X_hh"], 1e-11)
        _close(orc.lstsq_mgs_batched(A, B), golden["ls/cfg3/X_mgs"], 1e-11)
        assert golden["ls/cfg3/X_mgs"].shape == (2, 64 * 16)  # ravel quirk, qr.py:119
        return
    b = golden[f"ls/{name}/b"]
    # upper-triangular 50x50 systems with entries in [-100, 100] are ill-conditioned: 1e-9 here
    tol = 1e-9 if name.startswith("upper50") else 1e-11
    _close(orc.lstsq_householder(A, b), golden[f"ls/{name}/x_hh"], tol)
    _close(orc.lstsq_mgs(A, b), golden[f"ls/{name}/x_mgs"], tol)


@pytest.mark.parametrize("name", golden_cases("svd"))
def test_svd_matches_reference(golden, name):
    A = golden[f"svd/{name}/A"]
    rng = np.random.RandomState(999)
    U, s, Vt = orc.svd_gram(A, rng=rng)
    s_ref = golden[f"svd/{name}/s"]
    np.testing.assert_allclose(s, s_ref, rtol=1e-10, atol=1e-12)  # tests/test_svd.py:52
    Vt_ref = golden[f"svd/{name}/Vt"]
    r = int(np.sum(s_ref > 1e-8))
    # eigenvectors: up to a per-row sign (tests/test_svd.py:31-35), only for non-degenerate sigma
    sg = np.sign(np.sum(Vt[:r] * Vt_ref[:r], axis=1))
    assert orc.rel_max_err(Vt[:r] * sg[:, None], Vt_ref[:r]) <= 1e-8
    if f"svd/{name}/U" in golden:
        U_ref = golden[f"svd/{name}/U"]
        assert U.shape == U_ref.shape
        assert orc.rel_max_err(U[:, :r] * sg[None, :r], U_ref[:, :r]) <= 1e-8
        k = min(A.shape)
        assert np.linalg.norm((U[:, :k] * s[:k]) @ Vt[:k] - A) < 1e-10


def test_tsqr_convention_is_mgs(golden):
    A = golden["tsqr/mgs_1024x64/A"]
    Q, R = orc.tsqr_reference(A)
    _close(R, golden["tsqr/mgs_1024x64/R"], 1e-11)
    assert np.all(np.diag(R) > 0)


def test_error_cases():
    with pytest.raises(ValueError, match="linearly dependent"):
        orc.mgs_qr(np.ones((4, 2)))
    with pytest.raises(ValueError):
        orc.mgs_qr(np.ones((2, 3)))  # m < n always dependent
    with pytest.raises(ValueError):
        orc.householder_qr(np.ones((2, 3)))  # matmul shape mismatch upstream
    with pytest.raises(ValueError):
        orc.householder_qr(np.ones(3))


def test_householder_sign_convention():
    """R[j,j] = -copysign(||x||, x0) for EVERY column, also the last one of a square matrix."""
    A = np.random.default_rng(0).standard_normal((32, 32))
    Q, R = orc.householder_qr(A)
    Ql, Rl = np.linalg.qr(A)
    assert orc.rel_max_err(R[:31], Rl[:31]) < 1e-12
    assert orc.rel_max_err(R[31], -Rl[31]) < 1e-12  # LAPACK leaves the 1x1 tail alone
    assert orc.qr_residual(A, Q, R) < 1e-14 and orc.orth_error(Q) < 1e-13
