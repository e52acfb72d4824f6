"""Device plumbing shared by the host-side mirrors: moving numpy / torch inputs onto the
current CUDA device, the cached workspace, and the stream handle handed to the C ABI.
PyTorch is used for device memory and streams only; all arithmetic happens in
``libmriacl_recon.so``.
"""
from __future__ import annotations

import os
from typing import Any, Optional, Tuple

import numpy as np
import torch

from .adapters import recon_cabi

#: slices per launch group of the fused stage = slices the workspace can hold (4.4 MB of intermediate
#: per 15-coil knee slice at 4x); larger batches are cut into groups.  Override with MRIACL_CHUNK_SLICES
DEFAULT_CHUNK_SLICES = int(os.environ.get("MRIACL_CHUNK_SLICES", "64"))

_workspaces: dict = {}


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("mri_acl_imagesegmentation_adsp_b200 needs a CUDA device: the k-space -> image "
                           "stage runs only as sm_100a kernels (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def lib() -> recon_cabi.ReconLibrary:
    return recon_cabi.library()


def stream_ptr() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def workspace(nbytes: int) -> torch.Tensor:
    """Grow-only uint8 scratch per (device, stream); stream-ordered reuse is safe because every
    library call is enqueued on that same stream."""
    dev = torch.cuda.current_device()
    key = (dev, stream_ptr())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=torch.device("cuda", dev))
        _workspaces[key] = ws
    return ws


class Moved:
    """An input moved to the device together with how to hand results back."""

    def __init__(self, tensor: torch.Tensor, kind: str, home: Optional[torch.device], wide: bool = False):
        self.tensor, self.kind, self.home, self.wide = tensor, kind, home, wide

    def back(self, t: torch.Tensor, widen: bool = False) -> Any:
        """``widen``: the function is a twin of a reference function whose result dtype follows its input
        (complex128 in -> complex128 / float64 out, ``src/utils/kspace.py:4-20``); the arithmetic itself is
        single precision on the device."""
        if widen and self.wide:
            t = t.to(torch.complex128 if t.is_complex() else torch.float64)
        if self.kind == "numpy":
            return t.cpu().numpy()
        if self.home is not None and self.home.type == "cpu":
            return t.cpu()
        return t


def to_device_complex(x: Any, *, name: str = "kspace") -> Moved:
    """numpy complex / torch complex64 / torch float32 real view (..., 2), CPU or CUDA ->
    contiguous complex64 CUDA tensor (zero-copy when already there)."""
    dev = require_cuda()
    if isinstance(x, np.ndarray):
        if not np.iscomplexobj(x):
            raise ValueError(f"{name} must be complex, got dtype {x.dtype}")
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.complex64)).to(dev, non_blocking=True)
        return Moved(t, "numpy", None, wide=x.dtype == np.complex128)
    if isinstance(x, torch.Tensor):
        home = x.device
        wide = x.dtype == torch.complex128
        if not x.is_complex():
            if x.shape[-1] != 2:
                raise ValueError("Tensor does not have separate complex dim.")
            x = torch.view_as_complex(x.to(torch.float32).contiguous())
        t = x.to(device=dev, dtype=torch.complex64, non_blocking=True).contiguous()
        return Moved(t, "torch", home, wide=wide)
    raise ValueError(f"{name}: unsupported input type {type(x)!r}")


def to_device_real(x: Any, *, name: str = "data") -> Moved:
    dev = require_cuda()
    if isinstance(x, np.ndarray):
        if np.iscomplexobj(x):
            raise ValueError(f"{name} must be real, got dtype {x.dtype}")
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev, non_blocking=True)
        return Moved(t, "numpy", None)
    if isinstance(x, torch.Tensor):
        home = x.device
        if x.is_complex():
            raise ValueError(f"{name} must be real")
        return Moved(x.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous(), "torch", home)
    raise ValueError(f"{name}: unsupported input type {type(x)!r}")


def host_mask(mask: Any, width: int) -> Optional[np.ndarray]:
    """Sampling mask -> host float32 vector of length ``width`` (the C ABI takes it on the host and
    caches one device plan per distinct mask)."""
    if mask is None:
        return None
    if isinstance(mask, torch.Tensor):
        mask = mask.detach().cpu().numpy()
    m = np.asarray(mask, dtype=np.float32).reshape(-1)
    if m.shape[0] != width:
        raise ValueError(f"sampling mask has {m.shape[0]} entries, k-space width is {width}")
    return np.ascontiguousarray(m)


def batch_dims(shape: Tuple[int, ...], keep: int) -> Tuple[Tuple[int, ...], int]:
    lead = tuple(shape[:-keep]) if keep else tuple(shape)
    n = 1
    for d in lead:
        n *= int(d)
    return lead, n
