"""`engine="b200"` inside the UNMODIFIED reference: the three INTEGRATION.md edits applied as a runtime patch.

flox resolves an engine's kernels by name in `generic_aggregate` (flox/aggregations.py:60-133) and raises ValueError for
an unknown engine string (`:113-116`); `chunk_reduce` has ONE engine-specific pre-step, the sort for engine="flox"
(flox/core.py:336-339), which any other engine name already skips.  So the whole seam is the dispatch branch:

    flox_b200.plugin.install()                       # once per process
    flox.groupby_reduce(array, by, func="nanmean", engine="b200")
    flox.core.chunk_reduce(block, by_block, func=agg.chunk, ..., engine="b200")     # what a dask task runs

`install` wraps `generic_aggregate` (in every flox module that imported the name) with a function that serves
engine="b200" from `flox_b200.aggregate_b200` — same call signature, numpy in / numpy out — and hands every other
engine to the original.  Reductions the b200 engine does not provide fall back to the reference's in-repo numpy kernels
(`aggregate_flox`) exactly as engine="flox" falls back to numpy_groupies for names it lacks (`:85-90`).
Nothing in the reference is edited on disk; `uninstall` restores the originals.
"""
from __future__ import annotations

import functools
import sys

ENGINE = "b200"
_state: dict = {}


def install(flox=None):
    """Patch the imported reference package `flox` (default: sys.modules["flox"]).  Idempotent."""
    if flox is None:
        flox = sys.modules.get("flox")
    if flox is None or not hasattr(flox, "groupby_reduce"):
        raise RuntimeError("import the reference `flox` package first")
    if _state.get("flox") is flox:
        return flox
    import flox.aggregate_flox as aggregate_flox
    import flox.aggregations as aggregations

    from . import aggregate_b200

    original = aggregations.generic_aggregate

    @functools.wraps(original)
    def generic_aggregate(group_idx, array, *, engine, func, axis=-1, size=None, fill_value=None, dtype=None, **kwargs):
        if engine != ENGINE or func == "identity":
            return original(group_idx, array, engine=engine, func=func, axis=axis, size=size, fill_value=fill_value,
                            dtype=dtype, **kwargs)
        method = getattr(aggregate_b200, func, None) if isinstance(func, str) else None
        if method is None:  # not a scatter reduction (quantile, median, scans, ...): the reference's own numpy kernels,
            # which need what chunk_reduce only does for engine="flox" — rows sorted by group (core.py:336-339)
            if getattr(group_idx, "ndim", 0) == 1 and hasattr(aggregate_flox, str(func)):
                group_idx, array, _ = aggregate_flox._prepare_for_flox(group_idx, array)
            return original(group_idx, array, engine="flox", func=func, axis=axis, size=size, fill_value=fill_value,
                            dtype=dtype, **kwargs)
        return method(group_idx, array, axis=axis, size=size, fill_value=fill_value, dtype=dtype, **kwargs)

    patched = []
    for name, mod in list(sys.modules.items()):
        if name == "flox" or name.startswith("flox."):
            if getattr(mod, "generic_aggregate", None) is original:
                mod.generic_aggregate = generic_aggregate
                patched.append(mod)
    _state.update(flox=flox, original=original, patched=patched, aggregate_flox=aggregate_flox)
    return flox


def uninstall():
    for mod in _state.get("patched", []):
        mod.generic_aggregate = _state["original"]
    _state.clear()
