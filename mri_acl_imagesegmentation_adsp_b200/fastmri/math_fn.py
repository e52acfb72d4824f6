"""Twins of ``ZIP!/DL_reconstruction/math_fn.py:55-86`` on the real-view layout ``(..., 2)``."""
from __future__ import annotations

import torch

from .. import _device as D


def _abs(data: torch.Tensor, squared: bool) -> torch.Tensor:
    if not data.shape[-1] == 2:
        raise ValueError("Tensor does not have separate complex dim.")
    mv = D.to_device_complex(data, name="data")
    t = mv.tensor
    out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
    if t.numel():
        D.lib().complex_abs(t.data_ptr(), out.data_ptr(), t.numel(), squared, D.stream_ptr())
    return mv.back(out)


def complex_abs(data: torch.Tensor) -> torch.Tensor:
    return _abs(data, False)


def complex_abs_sq(data: torch.Tensor) -> torch.Tensor:
    return _abs(data, True)
