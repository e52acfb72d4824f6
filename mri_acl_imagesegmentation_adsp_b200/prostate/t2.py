"""Twins of the vendored prostate T2 reconstruction tail
(``ZIP!/fastmri_prostate/reconstruction/t2/prostate_t2_recon.py:65-121``,
``reconstruction/utils.py:7-29,54-73``, ``data/mri_data.py:63-85,123-160``).  GRAPPA (``:27-63``) is
out of scope."""
from __future__ import annotations

from typing import Any, Optional, Sequence, Tuple

import numpy as np
import torch

from ..recon.cartesian import zero_filled_rss
from ..utils import kspace as _k


def padding_lr(enc_x: int, max_pe_index: int) -> Tuple[int, int]:
    """``get_padding`` + the floor/ceil rule of ``zero_pad_kspace_hdr`` from the two header numbers
    (encodedSpace.matrixSize.x, encodingLimits.kspace_encoding_step_1.maximum): 640, 450 -> (94, 95)."""
    p = (enc_x - (max_pe_index + 1)) / 2
    if p % 2 != 0:
        return int(np.floor(p)), int(np.ceil(p))
    return int(p), int(p)


def ifftnd(kspace: Any, axes: Optional[Sequence[int]] = (-2, -1)) -> Any:
    """Centred orthonormal inverse FFT over the LAST TWO axes (the only use on the path,
    ``prostate_t2_recon.py:99``: ``ifftnd(data_sl, [1, 2])`` on ``(C, RO, PE)``)."""
    nd = kspace.ndim
    ax = sorted(a % nd for a in axes)
    if ax != [nd - 2, nd - 1]:
        raise ValueError("only the last two axes are supported")
    return _k.ifft2c(kspace)


def create_coil_combined_im(multicoil_multislice_kspace: Any) -> Any:
    """``(S, C, RO, PE)`` complex -> ``(S, RO, PE)`` float64: iFFT, RSS over coils, ``np.flipud``
    (``prostate_t2_recon.py:80-102``).  Computed in float32 on the device, widened on return."""
    img, _, _ = zero_filled_rss(multicoil_multislice_kspace, None, None, None, flip_rows=True)
    return img.astype(np.float64) if isinstance(img, np.ndarray) else img.to(torch.float64)


def t2_average_combine(kspace: Any, pad: Tuple[int, int], crop: Tuple[int, int] = (320, 320), mask: Any = None) -> Any:
    """``(A, S, C, RO, PE)`` -> ``(S, oh, ow)`` float64: zero-pad PE, coil-combine every average, mean
    over averages AFTER the RSS, centre crop (``prostate_t2_recon.py:65-75``) -- one fused call."""
    img, _, _ = zero_filled_rss(kspace, mask, crop, None, flip_rows=True, average_axis=0, pad=pad)
    return img.astype(np.float64) if isinstance(img, np.ndarray) else img.to(torch.float64)
