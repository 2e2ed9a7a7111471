"""NaN predicates used on the hot path (mirror of flox/xrutils.py:149-199)."""
from __future__ import annotations

import numpy as np


def isnull(data):
    """True where missing: NaN for floats/complex, NaT for datetimes, never for ints/bools."""
    if hasattr(data, "is_floating_point"):  # torch tensor
        import torch

        return torch.isnan(data) if data.is_floating_point() else torch.zeros_like(data, dtype=torch.bool)
    if np.isscalar(data) or data is None:
        import pandas as pd

        return pd.isnull(data)
    data = np.asarray(data)
    kind = data.dtype.kind
    if kind in "fc":
        return np.isnan(data)
    if kind in "mM":
        return np.isnat(data)
    if kind in "biu":
        return np.zeros(data.shape, dtype=bool)
    import pandas as pd

    return pd.isnull(data)


def notnull(data):
    return ~isnull(data)


def is_scalar(value) -> bool:
    return np.isscalar(value) or value is None or (hasattr(value, "ndim") and value.ndim == 0) or isinstance(value, (str, bytes))
