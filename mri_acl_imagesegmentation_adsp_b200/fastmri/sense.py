"""SENSE-style coil combine: a second coil-combine mode beside root-sum-of-squares.

``sens_combine(img, sens)`` = ``sum_c img_c * conj(sens_c)`` over the coil axis 1 -- ``np.sum(img * sens.conj(), axis=1)`` of
``ZIP!/fastmri_prostate/reconstruction/dwi/prostate_dwi_recon.py:106-108`` and ``sens_reduce`` of
``ZIP!/DL_reconstruction/models/varnet.py:199-203`` (there on real views, after the iFFT); ``magnitude=True`` adds the
``np.abs`` of ``:109``.  One C-ABI call (``mriacl_sense_combine``)."""
from __future__ import annotations

from typing import Any

import torch

from .. import _device as D


def sens_combine(img: Any, sens: Any, magnitude: bool = False) -> Any:
    """``img`` complex ``(B, C, ...)``; ``sens`` the same shape or ``(1, C, ...)`` / ``(C, ...)`` (shared maps).  Returns
    ``(B, ...)`` complex64, or float32 with ``magnitude``."""
    mv = D.to_device_complex(img, name="img")
    ms = D.to_device_complex(sens, name="sens")
    x, s = mv.tensor, ms.tensor
    if x.ndim < 3:
        raise ValueError(f"img must be (B, C, ...), got {tuple(x.shape)}")
    if s.ndim == x.ndim - 1:
        s = s[None]
    if s.shape[1:] != x.shape[1:] or s.shape[0] not in (1, x.shape[0]):
        raise ValueError(f"sens shape {tuple(s.shape)} does not match img shape {tuple(x.shape)}")
    b, c = x.shape[0], x.shape[1]
    n = 1
    for d in x.shape[2:]:
        n *= int(d)
    out = torch.empty((b,) + tuple(x.shape[2:]), dtype=torch.float32 if magnitude else torch.complex64, device=x.device)
    if b and n:
        D.lib().sense_combine(x.data_ptr(), s.data_ptr(), out.data_ptr(), b, c, n, s.shape[0] == 1 and b > 1, magnitude, D.stream_ptr())
    return mv.back(out, widen=True)
