"""Scatter results from the groups that were found to the groups that are expected.

Behavioural mirror of the dense path of flox/reindex.py (`reindex_` :160-216, `reindex_numpy`
:92-103, `ReindexStrategy` :53-89, `ReindexArrayType` :23-50).  The sparse COO variant is out of
scope.  Works on numpy arrays and torch tensors (a device-side gather for the latter).
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum, auto

import numpy as np
import pandas as pd
import torch

from . import xrdtypes
from .arrays import is_torch, torch_dtype
from .xrutils import isnull


class ReindexArrayType(Enum):
    AUTO = auto()
    NUMPY = auto()
    SPARSE_COO = auto()

    def is_same_type(self, other) -> bool:
        if self is ReindexArrayType.AUTO:
            return True
        if self is ReindexArrayType.NUMPY:
            return isinstance(other, np.ndarray)
        return False


@dataclass
class ReindexStrategy:
    """blockwise: reindex at the per-chunk step?  array_type: container to reindex to."""

    blockwise: bool | None
    array_type: ReindexArrayType = ReindexArrayType.AUTO

    def __post_init__(self):
        if self.blockwise is True and self.array_type not in (ReindexArrayType.AUTO, ReindexArrayType.NUMPY):
            raise ValueError("Setting reindex.blockwise=True not allowed for non-numpy array type.")
        if self.array_type is ReindexArrayType.SPARSE_COO:
            raise NotImplementedError("sparse reindexing is out of scope for the b200 engine")

    def set_blockwise_for_numpy(self):
        self.blockwise = True if self.blockwise is None else self.blockwise


def reindex_(array, from_, to, *, array_type: ReindexArrayType = ReindexArrayType.AUTO, fill_value=None, axis: int = -1,
             promote: bool = False):
    """Gather `array` (groups `from_` on the last axis) onto groups `to`; missing -> fill_value."""
    if not isinstance(to, pd.Index):
        if promote:
            to = pd.Index(to)
        else:
            raise ValueError("reindex requires a pandas.Index or promote=True")
    if to.ndim > 1:
        raise ValueError(f"Cannot reindex to a multidimensional array: {to}")
    if axis not in (-1, array.ndim - 1):
        raise NotImplementedError("reindex_ works along the last axis")
    torchy = is_torch(array)
    if array.shape[-1] == 0:  # every label was missing
        shape = tuple(array.shape[:-1]) + (len(to),)
        if torchy:
            return torch.full(shape, fill_value, dtype=array.dtype, device=array.device)
        return np.full(shape, fill_value, dtype=array.dtype)
    if not isinstance(from_, pd.Index):
        from_ = pd.Index(from_.cpu().numpy() if is_torch(from_) else from_)
    if from_.equals(to) and array_type.is_same_type(array):  # O(1) for two RangeIndex objects
        return array
    if from_.dtype.kind == "O" and isinstance(from_[0], tuple):
        raise NotImplementedError("Currently does not support reindexing with object arrays of tuples.")
    np_dtype = np.dtype(str(array.dtype).replace("torch.", "")) if torchy else array.dtype
    if fill_value == xrdtypes.NA or (fill_value is not None and np.all(isnull(fill_value))):
        new_dtype, fill_value = xrdtypes.maybe_promote(np_dtype)
    else:
        new_dtype = np_dtype
    idx = from_.get_indexer(to)
    missing = idx == -1
    if torchy:
        gather = torch.as_tensor(np.where(missing, 0, idx), device=array.device)
        out = array.index_select(array.ndim - 1, gather)
        if missing.any():
            if fill_value is None:
                raise ValueError("Filling is required. fill_value cannot be None.")
            out = out.to(torch_dtype(new_dtype))
            out[..., torch.as_tensor(missing, device=array.device)] = fill_value
        return out
    out = array[..., idx]
    if missing.any():
        if fill_value is None:
            raise ValueError("Filling is required. fill_value cannot be None.")
        out = out.astype(new_dtype, copy=False)
        out[..., missing] = fill_value
    return out
