"""The plug-in seam inside the UNMODIFIED reference (CPU tier: the kernel layer is the numpy emulator of
tests/emu_backend.py, everything above it — flox's own groupby_reduce / chunk_reduce / simple_combine /
_finalize_results and flox_b200's engine functions — is the real code)."""
import numpy as np
import pytest

from oracle import ref_loader
from tests import plugin_cases as P
from tests.cpu_fixtures import emu  # noqa: F401

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="needs /root/reference or baseline/_ref")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64])
def test_reference_groupby_reduce_with_engine_b200(dtype, emu):  # noqa: F811
    P.check_groupby_reduce(P.load_patched_reference(), dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32])
def test_reference_graph_blocks_with_engine_b200(dtype, emu):  # noqa: F811
    P.check_blocks_through_reference_graph(P.load_patched_reference(), dtype)


def test_unknown_engine_still_rejected_and_flox_untouched(emu):  # noqa: F811
    flox = P.load_patched_reference()
    x, by, G = P.inputs(np.float64)
    with pytest.raises(ValueError):
        flox.groupby_reduce(x, by, func="sum", engine="b300")
    a, _ = flox.groupby_reduce(x, by, func="sum", engine="flox")
    assert a.shape == (3, G - 1)


def test_non_scatter_reductions_fall_back_to_the_reference_kernels(emu):  # noqa: F811
    flox = P.load_patched_reference()
    x, by, G = P.inputs(np.float64)
    for func in ("nanmedian", "median"):
        want, _ = flox.groupby_reduce(x, by, func=func, engine="flox")
        got, _ = flox.groupby_reduce(x, by, func=func, engine="b200")
        np.testing.assert_array_equal(got, want)
