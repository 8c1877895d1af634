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
    m, n = 1 << 16, 128
    A = np.random.default_rng(6).standard_normal((m, n))
    lo, hi = d.my_row_slice(m, info)
    Al = np.ascontiguousarray(A[lo:hi])
    dA, dQ, dR = ctx.upload(Al), ctx.alloc(Al.nbytes), ctx.alloc(8 * n * n)
    ctx.call("lq_tsqr_sharded_dev", dA.ptr, hi - lo, n, dQ.ptr, dR.ptr)
    Q, R = ctx.download(dQ, (hi - lo, n)), ctx.download(dR, (n, n))
    dU, ds, dVt = ctx.alloc(Al.nbytes), ctx.alloc(8 * n), ctx.alloc(8 * n * n)
    rk = C.c_int(0)
    ctx.call("lq_svd_gram_sharded_dev", dA.ptr, hi - lo, n, C.c_double(1e-12), dU.ptr, ds.ptr, dVt.ptr, C.byref(rk))
    U, s, Vt = ctx.download(dU, (hi - lo, n)), ctx.download(ds, (n,)), ctx.download(dVt, (n, n))
    np.savez(os.path.join(outdir, f"r{rank}.npz"), lo=lo, hi=hi, Q=Q, R=R, U=U, s=s, Vt=Vt, rank=rk.value)
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
    m, n = 1 << 16, 128
    A = np.random.default_rng(6).standard_normal((m, n))
    z = [np.load(tmp_path / f"r{r}.npz") for r in range(2)]
    assert np.array_equal(z[0]["R"], z[1]["R"]) and np.array_equal(z[0]["s"], z[1]["s"])  # replicated, bitwise
    Q = np.vstack([z[0]["Q"], z[1]["Q"]])
    U = np.vstack([z[0]["U"], z[1]["U"]])
    R, sv, Vt = z[0]["R"], z[0]["s"], z[0]["Vt"]
    Qo, Ro = orc.tsqr_reference(A)
    assert orc.rel_max_err(R, Ro) <= 1e-10 and orc.orth_error(Q) <= 1e-12 and orc.qr_residual(A, Q, R) <= 1e-12
    _, so, _ = orc.svd_gram(A)
    assert np.max(np.abs(sv - so) / so) <= 1e-10
    assert np.linalg.norm((U * sv) @ Vt - A) / np.linalg.norm(A) <= 1e-12
