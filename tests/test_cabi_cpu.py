"""The C-ABI library builds, loads and exports every symbol include/flox_b200.h declares
(no compute without a GPU), and the package refuses to run without CUDA instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

from flox_b200 import _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared():
    text = open(os.path.join(ROOT, "include", "flox_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(floxb_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from flox_b200 import build

        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/flox_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES out of sync with the header"
    assert lib.floxb_version() >= 100
    lib.floxb_strerror.restype = ctypes.c_char_p
    assert lib.floxb_strerror(-2) == b"unsupported dtype combination"


def test_argument_validation_without_gpu():
    lib = _lib.load()
    # invalid arguments are rejected before any CUDA call
    assert lib.floxb_moments(None, -1, 0, 3, 1, 1, 4, 0, None, None, None, 1, 0, None, 0, None) == -1
    assert lib.floxb_moments_workspace_bytes(_lib.F32, 1, 1, 100) == 0  # small G: replicated slots, no workspace
    assert lib.floxb_moments_workspace_bytes(_lib.F64, 1, 1, 10_000) == 0
    assert lib.floxb_combine(0, 0, None, 0, 4, None, None) == -1
    assert lib.floxb_smem_group_limit(_lib.F32, 0) >= 10_000


@pytest.mark.parametrize("src", ["int64", "int32"])
@pytest.mark.parametrize("bits", [16, 32])
def test_host_pack_codes(src, bits):
    """floxb_host_pack_codes is a HOST function (label staging, include/flox_b200.h): out-of-range and negative
    codes become -1 — exactly the rows the reduction kernels skip — everything else is narrowed unchanged."""
    import numpy as np

    lib = _lib.load()
    rng = np.random.default_rng(3)
    G = 30_000
    c = rng.integers(-50, G + 50, 100_003).astype(src)
    c[:6] = [0, G - 1, G, -1, np.iinfo(src).max, np.iinfo(src).min]
    out = np.full(c.size, 7, dtype=f"int{bits}")
    rc = lib.floxb_host_pack_codes(c.ctypes.data, _lib.I64 if src == "int64" else _lib.I32, c.size, G, out.ctypes.data, bits)
    assert rc == 0
    want = np.where((c >= 0) & (c < G), c, -1).astype(out.dtype)
    np.testing.assert_array_equal(out, want)
    assert lib.floxb_host_pack_codes(c.ctypes.data, _lib.I64, c.size, 40_000, out.ctypes.data, 16) == -4 or bits == 32
    assert lib.floxb_host_pack_codes(c.ctypes.data, _lib.F32, c.size, G, out.ctypes.data, bits) == -2


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine WITHOUT a GPU")
def test_no_cpu_fallback():
    import numpy as np

    import flox_b200
    from flox_b200 import kernels

    kernels.set_backend(None)
    with pytest.raises(_lib.FloxB200Error):
        flox_b200.groupby_reduce(np.ones(4), np.array([0, 0, 1, 1]), func="sum")
