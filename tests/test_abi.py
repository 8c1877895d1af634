"""The C-ABI library loads on a CPU-only box and exports every symbol include/linalg_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "linalg_b200.h")


def header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(lq_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as ge

    ge.build()
    from linalg_b200 import _native

    assert os.path.exists(_native.LIB_PATH)
    return _native.LIB_PATH


def test_header_lists_functions():
    names = header_functions()
    assert len(names) >= 40
    for must in ("lq_householder_qr_batched", "lq_mgs_qr_batched", "lq_lstsq_householder_batched", "lq_svd_gram", "lq_tsqr"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in header_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_python_signatures_match_header(lib_path):
    from linalg_b200 import _native

    assert sorted(_native.SIGNATURES) == header_functions()
    lib = _native.load_library()
    assert lib.lq_version().startswith(b"linalg_b200")


def test_library_is_sm100a_only(lib_path):
    out = subprocess.run(["cuobjdump", "--list-elf", lib_path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_torch_or_oracle_in_product():
    """The product never imports torch at module scope and never touches oracle/ (no CPU fallback)."""
    pkg = os.path.join(ROOT, "linalg_b200")
    for f in os.listdir(pkg):
        if not f.endswith(".py"):
            continue
        src = open(os.path.join(pkg, f)).read()
        assert "oracle" not in src.replace("no CPU fallback", ""), f
        for line in src.splitlines():
            if re.match(r"^(import torch|from torch)", line):
                raise AssertionError(f"{f}: module-scope torch import")


def test_no_device_raises_loudly(lib_path):
    """Without a GPU, creating a context raises; nothing silently falls back to the CPU."""
    from linalg_b200 import _native

    lib = _native.load_library()
    n = ctypes.c_int(-1)
    rc = lib.lq_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.Context(0)
