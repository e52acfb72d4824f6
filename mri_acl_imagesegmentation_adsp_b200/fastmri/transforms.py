"""Twins of ``ZIP!/DL_reconstruction/data/transforms.py``: to_tensor, center_crop, complex_center_crop,
center_crop_to_smallest, normalize, normalize_instance."""
from __future__ import annotations

from typing import Tuple, Union

import numpy as np
import torch

from .. import _device as D


def to_tensor(data: np.ndarray) -> torch.Tensor:
    """complex ndarray -> real view ``(..., 2)`` tensor (``transforms.py:14-29``); host-side relayout."""
    if np.iscomplexobj(data):
        data = np.stack((data.real, data.imag), axis=-1)
    return torch.from_numpy(data)


def center_crop(data: torch.Tensor, shape: Tuple[int, int]) -> torch.Tensor:
    """Centre crop of the last two axes (``transforms.py:45-67``); ValueError("Invalid shapes.") when
    the crop exceeds the data.  A view, exactly like the reference (no kernel needed)."""
    if not (0 < shape[0] <= data.shape[-2] and 0 < shape[1] <= data.shape[-1]):
        raise ValueError("Invalid shapes.")
    w_from = (data.shape[-2] - shape[0]) // 2
    h_from = (data.shape[-1] - shape[1]) // 2
    return data[..., w_from:w_from + shape[0], h_from:h_from + shape[1]]


def complex_center_crop(data: torch.Tensor, shape: Tuple[int, int]) -> torch.Tensor:
    """Centre crop of dims -3 and -2 of a real-view complex tensor ``(..., H, W, 2)`` (``transforms.py:70-92``);
    a view, ValueError("Invalid shapes.") when the crop exceeds the data."""
    if not (0 < shape[0] <= data.shape[-3] and 0 < shape[1] <= data.shape[-2]):
        raise ValueError("Invalid shapes.")
    w_from = (data.shape[-3] - shape[0]) // 2
    h_from = (data.shape[-2] - shape[1]) // 2
    return data[..., w_from:w_from + shape[0], h_from:h_from + shape[1], :]


def center_crop_to_smallest(x: torch.Tensor, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Crop both images to the smaller extent of each of the last two dims (``transforms.py:95-117``)."""
    smallest_width = min(x.shape[-1], y.shape[-1])
    smallest_height = min(x.shape[-2], y.shape[-2])
    return center_crop(x, (smallest_height, smallest_width)), center_crop(y, (smallest_height, smallest_width))


def normalize(data, mean, stddev, eps=0.0):
    """``(data - mean) / (stddev + eps)`` (``transforms.py:120-140``)."""
    return (data - mean) / (stddev + eps)


def normalize_instance(data: torch.Tensor, eps: Union[float, torch.Tensor] = 0.0
                       ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Instance normalisation of ONE tensor with the unbiased std (``transforms.py:143-162``).
    Returns ``(normalised, mean, std)`` with 0-d mean/std."""
    mv = D.to_device_real(data)
    t = mv.tensor
    out = torch.empty_like(t)
    ms = torch.empty((1, 2), dtype=torch.float32, device=t.device)
    D.lib().normalize_instance(t.data_ptr(), out.data_ptr(), ms.data_ptr(), 1, t.numel(), float(eps), D.stream_ptr())
    return mv.back(out), mv.back(ms[0, 0]), mv.back(ms[0, 1])
