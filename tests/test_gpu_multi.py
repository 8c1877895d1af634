"""Row-sharded TSQR / Gram-SVD over NCCL with 2 ranks (needs >= 2 GPUs; skipped on a 1-GPU box)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import ctypes

    from linalg_b200 import _native

    n = ctypes.c_int(0)
    _native.load_library().lq_device_count(ctypes.byref(n))
    return n.value


CASES = {
    # name: (rows, householder tree forced, column scaling exponent range, uneven split)
    "gaussian": (1 << 16, False, 0.0, False),
    "householder_tree": (1 << 14, True, 0.0, False),          # LINALG_B200_TSQR_HOUSEHOLDER: all-gather of the R factors
    "ill_conditioned": (1 << 14, False, 9.0, False),           # cond ~ 1e9: CholeskyQR2 must hand over to the reflector path
    "uneven_shards": ((1 << 13) + 517, False, 0.0, True),      # ranks of different height take the same (collective) decision
}


def _case_matrix(name):
    m, _, span, _ = CASES[name]
    n = 128
    A = np.random.default_rng(6).standard_normal((m, n))
    if span:
        A = A * np.logspace(0.0, -span, n)[None, :]
    return A


def _case_bounds(name, rank, world):
    m, _, _, uneven = CASES[name]
    if uneven:  # rank 0 gets 3/8 of the rows (4n - 1 would be the edge of the advisor's finding; keep both tall)
        cut = (3 * m) // 8
        return (0, cut) if rank == 0 else (cut, m)
    from linalg_b200.utils import shard_bounds

    return shard_bounds(m, world, rank)


def _worker(rank, world, port, outdir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), LINALG_B200_DEVICE=str(rank))
    sys.path.insert(0, ROOT)
    import ctypes as C

    import linalg_b200 as lb
    from linalg_b200 import dist as d

    info = d.init_control_plane("gloo")
    ctx = lb.Context(rank)
    d.init_comm(ctx, info)
    n = 128
    for name, (m, tree, _, _) in CASES.items():
        A = _case_matrix(name)
        lo, hi = _case_bounds(name, rank, world)
        Al = np.ascontiguousarray(A[lo:hi])
        ctx.set_option("TSQR_HOUSEHOLDER", tree)
        dA, dQ, dR = ctx.upload(Al), ctx.alloc(Al.nbytes), ctx.alloc(8 * n * n)
        ctx.call("lq_tsqr_sharded_dev", dA.ptr, hi - lo, n, dQ.ptr, dR.ptr)
        Q, R = ctx.download(dQ, (hi - lo, n)), ctx.download(dR, (n, n))
        ctx.set_option("TSQR_HOUSEHOLDER", False)
        dU, ds, dVt = ctx.alloc(Al.nbytes), ctx.alloc(8 * n), ctx.alloc(8 * n * n)
        rk = C.c_int(0)
        ctx.call("lq_svd_gram_sharded_dev", dA.ptr, hi - lo, n, C.c_double(1e-12), dU.ptr, ds.ptr, dVt.ptr, C.byref(rk))
        U, s, Vt = ctx.download(dU, (hi - lo, n)), ctx.download(ds, (n,)), ctx.download(dVt, (n, n))
        np.savez(os.path.join(outdir, f"{name}_r{rank}.npz"), lo=lo, hi=hi, Q=Q, R=R, U=U, s=s, Vt=Vt, rank=rk.value)
        for b in (dA, dQ, dR, dU, ds, dVt):
            b.free()
    d.barrier()
    ctx.call("lq_comm_destroy")
    d.shutdown_control_plane()


def test_sharded_tsqr_and_svd_two_ranks(tmp_path):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from oracle import linalg_oracle as orc

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    n = 128
    for name in CASES:
        A = _case_matrix(name)
        z = [np.load(tmp_path / f"{name}_r{r}.npz") for r in range(2)]
        assert int(z[0]["lo"]) == 0 and int(z[0]["hi"]) == int(z[1]["lo"]) and int(z[1]["hi"]) == A.shape[0], name
        assert np.array_equal(z[0]["R"], z[1]["R"]) and np.array_equal(z[0]["s"], z[1]["s"]), name  # replicated, bitwise
        assert np.array_equal(z[0]["Vt"], z[1]["Vt"]), name
        Q = np.vstack([z[0]["Q"], z[1]["Q"]])
        U = np.vstack([z[0]["U"], z[1]["U"]])
        R, sv, Vt = z[0]["R"], z[0]["s"], z[0]["Vt"]
        # thin QR, diag(R) > 0 = the convention of linalg/qr.py:39-42 (reference MGS `qr`)
        assert np.all(np.diag(R) > 0) and np.max(np.abs(np.tril(R, -1))) == 0.0, name
        assert orc.orth_error(Q) <= 1e-12 and orc.qr_residual(A, Q, R) <= 1e-12, (name, orc.orth_error(Q), orc.qr_residual(A, Q, R))
        if name != "ill_conditioned":
            Qo, Ro = orc.tsqr_reference(A)
            assert orc.rel_max_err(R, Ro) <= 1e-10, (name, orc.rel_max_err(R, Ro))
            _, so, _ = orc.svd_gram(A)
            assert np.max(np.abs(sv - so) / so) <= 1e-10, name
            assert np.linalg.norm((U * sv) @ Vt - A) / np.linalg.norm(A) <= 1e-12, name
        else:
            # the reference's own MGS loses orthogonality at this conditioning; R is pinned by LAPACK up to row signs,
            # columnwise relative to the column norms (the scaling spans 9 decades)
            Rl = np.linalg.qr(A, mode="r")
            Rl = Rl * np.sign(np.diag(Rl))[:, None]
            cn = np.linalg.norm(A, axis=0)
            assert np.max(np.abs(R - Rl) / cn[None, :]) <= 1e-10, name
            # the A^T A route squares the condition number (the reference loses the small singular values the same way,
            # linalg/svd.py:42-54): only the leading ones are comparable
            so = np.linalg.svd(A, compute_uv=False)
            lead = so > 1e-4 * so[0]
            assert np.max(np.abs(sv[lead] - so[lead]) / so[lead]) <= 1e-6, name


def test_devices_kwarg_matches_one_gpu_bitwise():
    """householder_qr_batched / qr_batched / least_squares_*_batched(devices=[0, 1]): one process, one host thread and one
    context per GPU, contiguous batch split, no communication -- bitwise the single-GPU result (SURVEY.md sections 5, 8b)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import linalg_b200 as lb

    rng = np.random.default_rng(77)
    A = rng.standard_normal((4099, 32, 32))
    Q1, R1 = lb.householder_qr_batched(A)
    Q2, R2 = lb.householder_qr_batched(A, devices=[0, 1])
    assert np.array_equal(Q1, Q2) and np.array_equal(R1, R2)
    Q1, R1 = lb.qr_batched(A)
    Q2, R2 = lb.qr_batched(A, devices=[1, 0])
    assert np.array_equal(Q1, Q2) and np.array_equal(R1, R2)
    A3 = rng.standard_normal((301, 256, 64))
    B3 = rng.standard_normal((301, 256, 16))
    X1 = lb.least_squares_householder_qr_batched(A3, B3)
    X2 = lb.least_squares_householder_qr_batched(A3, B3, devices=[0, 1])
    assert np.array_equal(X1, X2)
    assert np.array_equal(lb.least_squares_qr_batched(A3, B3), lb.least_squares_qr_batched(A3, B3, devices=[0, 1]))
    # the reference's error surfaces from whichever device meets it
    A[4000, :, 7] = A[4000, :, 3]
    with pytest.raises(ValueError, match="linearly dependent"):
        lb.qr_batched(A, devices=[0, 1])
    with pytest.raises(ValueError):
        lb.householder_qr_batched(A, devices=[0, 0])
