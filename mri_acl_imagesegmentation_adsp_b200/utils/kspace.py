"""Drop-in for the reference's ``src/utils/kspace.py`` (same four names, argument meaning and
shapes), computed by ``libmriacl_recon.so`` on the current CUDA device.

numpy in -> numpy out exactly like the reference; torch tensors (CPU or CUDA) are accepted too
and come back as torch tensors on their original device.  Leading dimensions are batch.

dtype follows the input exactly as in the reference: complex64 -> complex64 / float32, complex128 ->
complex128 / float64.  The device arithmetic is single precision either way (the path's parity bar is
rel-L2 <= 1e-5), so a complex128 result carries float32 accuracy.
"""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from .. import _device as D


def _fft2c(x: Any, inverse: bool) -> Any:
    mv = D.to_device_complex(x, name="x")
    t = mv.tensor
    if t.ndim < 2:
        raise ValueError(f"need at least 2 dims, got {tuple(t.shape)}")
    h, w = t.shape[-2:]
    _, b = D.batch_dims(t.shape, 2)
    out = torch.empty_like(t)
    if t.numel():
        D.lib().fft2c(t.data_ptr(), out.data_ptr(), b, h, w, inverse, D.stream_ptr())
    return mv.back(out, widen=True)      # complex128 in -> complex128 out, as numpy.fft (kspace.py:7,14)


def fft2c(x: Any) -> Any:
    """Centred orthonormal 2-D FFT over the last two axes (``src/utils/kspace.py:4-9``)."""
    return _fft2c(x, False)


def ifft2c(x: Any) -> Any:
    """Centred orthonormal 2-D inverse FFT over the last two axes (``src/utils/kspace.py:11-16``)."""
    return _fft2c(x, True)


def complex_abs(x: Any) -> Any:
    """``sqrt(re^2 + im^2)`` (``src/utils/kspace.py:18-20``)."""
    mv = D.to_device_complex(x, name="x")
    t = mv.tensor
    out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
    if t.numel():
        D.lib().complex_abs(t.data_ptr(), out.data_ptr(), t.numel(), False, D.stream_ptr())
    return mv.back(out, widen=True)


def center_crop_or_pad(img: Any, out_h: int, out_w: int) -> Any:
    """Centre crop or zero-pad the last two axes to ``(out_h, out_w)`` (``src/utils/kspace.py:22-31``).
    Pure indexing: bit-exact.  float32 and complex64 payloads are supported."""
    is_complex = np.iscomplexobj(img) if isinstance(img, np.ndarray) else (isinstance(img, torch.Tensor) and img.is_complex())
    mv = D.to_device_complex(img, name="img") if is_complex else D.to_device_real(img, name="img")
    t = mv.tensor
    if t.ndim < 2:
        raise ValueError(f"need at least 2 dims, got {tuple(t.shape)}")
    h, w = t.shape[-2:]
    lead, b = D.batch_dims(t.shape, 2)
    out = torch.empty(lead + (int(out_h), int(out_w)), dtype=t.dtype, device=t.device)
    if out.numel():
        D.lib().center_crop_or_pad(t.data_ptr(), out.data_ptr(), b, h, w, int(out_h), int(out_w),
                                   8 if is_complex else 4, D.stream_ptr())
    return mv.back(out)
