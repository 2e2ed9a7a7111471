"""Install the UNMODIFIED reference (xarray-contrib/flox @ /root/reference) into baseline/_ref.

TEST / BENCH INFRASTRUCTURE ONLY.  `baseline/_ref/` is git-ignored (no reference source enters
the history) but not gpurun-ignored, so the installed package travels to the GPU box, where
`bench.py --impl reference` and the plug-in seam tests import it (there is no /root/reference
there).  The recipe is the one the task contract names:

    python -m pip install --no-index --no-build-isolation --no-deps \
        --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>

with two accommodations, both recorded in DESIGN.md §3:
  * the install runs from a copy under /tmp (the build writes an egg-info into the source tree and
    /root/reference is read-only);
  * flox's hard dependencies `numpy_groupies` and `toolz` are in neither this image nor
    /opt/wheelhouse, and its metadata step imports the package: the two import-only stand-ins of
    oracle/refshim/ (OUR files: a stub numpy_groupies whose attributes raise on use, and the five
    toolz helpers flox imports) are put on PYTHONPATH for the build and copied beside the
    installed package.  engine="flox" (the reference's in-repo numpy kernels) needs neither.

Run:  python oracle/install_reference.py [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE_ROOT = "/root/reference"
TARGET = os.path.join(ROOT, "baseline", "_ref")
SHIMS = os.path.join(HERE, "refshim")


def installed() -> bool:
    return os.path.isfile(os.path.join(TARGET, "flox", "core.py")) and os.path.isdir(os.path.join(TARGET, "toolz"))


def install(force: bool = False, verbose: bool = False) -> str | None:
    """Returns the install directory, or None when /root/reference is absent (GPU box: prebuilt only)."""
    if installed() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "flox")):
        return TARGET if installed() else None
    if os.path.isdir(TARGET):
        shutil.rmtree(TARGET)
    os.makedirs(TARGET, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="floxref_") as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REFERENCE_ROOT, src, ignore=shutil.ignore_patterns(".git", "__pycache__"))
        env = dict(os.environ)
        env["PYTHONPATH"] = SHIMS + os.pathsep + env.get("PYTHONPATH", "")
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True, env=env, stdout=None if verbose else subprocess.DEVNULL)
    for shim in ("numpy_groupies", "toolz"):
        shutil.copytree(os.path.join(SHIMS, shim), os.path.join(TARGET, shim), ignore=shutil.ignore_patterns("__pycache__"))
    if not installed():
        raise RuntimeError("reference install did not produce baseline/_ref/flox")
    return TARGET


if __name__ == "__main__":
    print(install(force="--force" in sys.argv, verbose=True))
