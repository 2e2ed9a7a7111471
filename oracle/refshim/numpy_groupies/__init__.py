"""Import-only stand-in for numpy_groupies (absent from this image).

TEST INFRASTRUCTURE ONLY. It lets `import flox` succeed from /root/reference so
that engine="flox" (the in-repo sort + reduceat kernels) can be run to generate
and pin golden vectors. Any attempt to actually *use* numpy_groupies raises.
"""


class _Unavailable:
    def __init__(self, name):
        self._name = name

    def __getattr__(self, item):
        raise ImportError(f"numpy_groupies.{self._name}.{item} is not available (import-only shim)")


aggregate_numpy = _Unavailable("aggregate_numpy")
aggregate_numba = _Unavailable("aggregate_numba")
__version__ = "0.0.0+shim"
