"""Import-only stand-in (see package docstring)."""


def aggregate(*args, **kwargs):
    raise ImportError("numpy_groupies is not installed; this is an import-only shim")
