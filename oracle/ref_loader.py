"""Load the UNMODIFIED reference (flox @ /root/reference) for oracle pinning.

TEST INFRASTRUCTURE ONLY — used by tests/golden/make_golden.py and by the
CPU-side pinning tests when /root/reference exists (build container only; the
GPU box has no /root/reference and nothing run there may import this).
"""
from __future__ import annotations

import os
import sys

REFERENCE_ROOT = "/root/reference"
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "flox"))


def load_reference():
    """Return the reference `flox` package imported from /root/reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    for p in (REFERENCE_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.append(p)  # append: the reference tree has its own `tests` package
    import flox  # noqa: F401
    import flox.core  # noqa: F401

    return flox
