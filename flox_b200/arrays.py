"""Container plumbing: numpy / torch in, torch-on-device inside, same kind out."""
from __future__ import annotations

import warnings

import numpy as np
import torch

from .kernels import default_device

# read-only inputs (broadcast views, arrays handed over by xarray) are only ever READ here; torch.from_numpy warns
# about them once per process
warnings.filterwarnings("ignore", message="The given NumPy array is not writable", category=UserWarning)

_NP2TORCH = {
    np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64, np.dtype("float16"): torch.float16,
    np.dtype("int8"): torch.int8, np.dtype("int16"): torch.int16, np.dtype("int32"): torch.int32,
    np.dtype("int64"): torch.int64, np.dtype("uint8"): torch.uint8, np.dtype("bool"): torch.bool,
}
_TORCH2NP = {v: k for k, v in _NP2TORCH.items()}
_TORCH2NP[torch.bfloat16] = np.dtype("float32")

# dtypes the kernels read directly; everything else is widened once on the way in
_KERNEL_DTYPES = (torch.float32, torch.float64, torch.int32, torch.int64)


class TransferStats:
    """Bytes moved host -> device by this package (bench.py reports them as h2d_bytes_per_step)."""
    h2d_bytes = 0


def count_h2d(t: torch.Tensor) -> None:
    TransferStats.h2d_bytes += t.numel() * t.element_size()


def is_torch(x) -> bool:
    return isinstance(x, torch.Tensor)


def np_dtype_of(x) -> np.dtype:
    if is_torch(x):
        return _TORCH2NP[x.dtype]
    return np.asarray(x).dtype if not hasattr(x, "dtype") else np.dtype(x.dtype)


def torch_dtype(dt) -> torch.dtype:
    if isinstance(dt, torch.dtype):
        return dt
    dt = np.dtype(dt)
    if dt.kind in "mM":
        return torch.int64
    if dt in _NP2TORCH:
        return _NP2TORCH[dt]
    if dt.kind == "u":  # uint16/32/64: widen (torch has limited unsigned support)
        return torch.int64
    raise TypeError(f"unsupported dtype {dt}")


def kernel_dtype(dt: torch.dtype) -> torch.dtype:
    """The dtype the CUDA kernels will read for an input of dtype `dt`."""
    if dt in _KERNEL_DTYPES:
        return dt
    if dt in (torch.float16, torch.bfloat16):
        return torch.float32
    if dt in (torch.int8, torch.int16, torch.uint8, torch.bool):
        return torch.int32
    return torch.int64


def to_device(x, device: torch.device | None = None, *, dtype: torch.dtype | None = None) -> torch.Tensor:
    """numpy / torch (any device) -> contiguous torch tensor on `device`."""
    device = device or default_device()
    if is_torch(x):
        t = x
    else:
        a = np.asarray(x)
        if a.dtype.kind in "mM":
            a = a.view(np.int64)
        elif a.dtype.kind == "u" and a.dtype.itemsize > 1:
            a = a.astype(np.int64)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.device != device and device.type == "cuda" and t.device.type == "cpu":
        count_h2d(t)
    if dtype is not None and t.dtype != dtype:
        t = t.to(device=device, dtype=dtype, non_blocking=True)
    elif t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def from_device(t: torch.Tensor, like, np_dtype: np.dtype | None = None):
    """Return `t` in the container kind of `like` (numpy array or torch tensor)."""
    if is_torch(like):
        if np_dtype is not None and np.dtype(np_dtype) in _NP2TORCH and t.dtype != _NP2TORCH[np.dtype(np_dtype)]:
            t = t.to(_NP2TORCH[np.dtype(np_dtype)])
        return t
    a = t.cpu().numpy()
    if np_dtype is not None and a.dtype != np.dtype(np_dtype):
        a = a.astype(np_dtype)
    return a
