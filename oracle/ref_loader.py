"""Load the UNMODIFIED reference (xarray-contrib/flox) for oracle pinning, the plug-in seam tests and
the CPU baseline of bench.py.

TEST / BENCH INFRASTRUCTURE ONLY — nothing under flox_b200/ imports this.

Two locations, same code:
  * /root/reference            the read-only source tree (build container only);
  * <repo>/baseline/_ref       the pip-installed copy made by oracle/install_reference.py
                               (git-ignored, travels to the GPU box with the gpurun snapshot).
`prefer="installed"` picks baseline/_ref first (what runs on the GPU box), the default picks the
source tree when it exists.  Both need the two import-only stand-ins for `numpy_groupies` and
`toolz` (oracle/refshim/, copied into baseline/_ref by the installer).
"""
from __future__ import annotations

import os
import sys

REFERENCE_ROOT = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "refshim")
INSTALLED_ROOT = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")


def _source_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "flox"))


def installed_available() -> bool:
    return os.path.isfile(os.path.join(INSTALLED_ROOT, "flox", "core.py"))


def reference_available() -> bool:
    return _source_available() or installed_available()


def reference_location(prefer: str = "source") -> str | None:
    order = (("installed", "source") if prefer == "installed" else ("source", "installed"))
    for kind in order:
        if kind == "source" and _source_available():
            return REFERENCE_ROOT
        if kind == "installed" and installed_available():
            return INSTALLED_ROOT
    return None


def load_reference(prefer: str = "source"):
    """Return the reference `flox` package (unmodified)."""
    if "flox" in sys.modules and hasattr(sys.modules["flox"], "groupby_reduce"):
        return sys.modules["flox"]
    where = reference_location(prefer)
    if where is None:
        raise RuntimeError("the reference is neither at /root/reference nor installed under baseline/_ref "
                           "(run `python oracle/install_reference.py` in the build container)")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    paths = (where, _SHIMS) if where == REFERENCE_ROOT else (where,)
    for p in paths:
        if p not in sys.path:
            sys.path.append(p)  # append: the reference tree has its own `tests` package
    import flox  # noqa: F401
    import flox.core  # noqa: F401

    return flox
