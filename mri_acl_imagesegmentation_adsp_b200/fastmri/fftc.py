"""Twins of the vendored fastMRI transforms ``ZIP!/DL_reconstruction/fftc.py`` on the real-view
layout ``(..., H, W, 2)`` (byte-identical to complex64 ``(..., H, W)``)."""
from __future__ import annotations

from typing import Any

import torch

from ..utils import kspace as _k


def _check(data: Any) -> None:
    if not data.shape[-1] == 2:
        raise ValueError("Tensor does not have separate complex dim.")   # fftc.py:27-28,54-55


def fft2c_new(data: torch.Tensor, norm: str = "ortho") -> torch.Tensor:
    """Centred 2-D FFT (``fftc.py:14-38``); only ``norm="ortho"`` (the reference's default) is built."""
    _check(data)
    if norm != "ortho":
        raise ValueError("only norm='ortho' is supported")
    out = _k.fft2c(data)
    return torch.view_as_real(out) if isinstance(out, torch.Tensor) else out


def ifft2c_new(data: torch.Tensor, norm: str = "ortho") -> torch.Tensor:
    """Centred 2-D inverse FFT (``fftc.py:41-65``)."""
    _check(data)
    if norm != "ortho":
        raise ValueError("only norm='ortho' is supported")
    out = _k.ifft2c(data)
    return torch.view_as_real(out) if isinstance(out, torch.Tensor) else out
