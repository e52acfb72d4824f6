"""Twins of ``ZIP!/DL_reconstruction/coil_combine.py``: root-sum-of-squares over the coil axis."""
from __future__ import annotations

import torch

from .. import _device as D


def _rss(t: torch.Tensor, dim: int, is_complex: bool) -> torch.Tensor:
    nd = t.ndim
    dim = dim % nd
    outer = 1
    for d in t.shape[:dim]:
        outer *= int(d)
    inner = 1
    for d in t.shape[dim + 1:]:
        inner *= int(d)
    out = torch.empty(t.shape[:dim] + t.shape[dim + 1:], dtype=torch.float32, device=t.device)
    if out.numel():
        D.lib().rss(t.data_ptr(), out.data_ptr(), outer, int(t.shape[dim]), inner, is_complex, D.stream_ptr())
    return out


def rss(data: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """``sqrt((data**2).sum(dim))`` for real data (``coil_combine.py:12-25``)."""
    mv = D.to_device_real(data)
    return mv.back(_rss(mv.tensor, dim, False))


def rss_complex(data: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """RSS of real-view complex data ``(..., 2)`` (``coil_combine.py:28-41``); ``dim`` counts the
    dims of the real view, as in the reference."""
    if not data.shape[-1] == 2:
        raise ValueError("Tensor does not have separate complex dim.")
    nd = data.ndim
    dim = dim % nd
    if dim == nd - 1:
        raise ValueError("dim must not be the complex dim")
    mv = D.to_device_complex(data)
    return mv.back(_rss(mv.tensor, dim, True))
