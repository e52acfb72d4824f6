"""Twins of the vendored prostate T2 reconstruction tail
(``ZIP!/fastmri_prostate/reconstruction/t2/prostate_t2_recon.py:65-121``,
``reconstruction/utils.py:7-29,54-73``, ``data/mri_data.py:63-85,123-160``).  GRAPPA (``:27-63``) is
out of scope."""
from __future__ import annotations

from typing import Any, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _device as D
from ..fastmri import coil_combine as _cc
from ..recon.cartesian import zero_filled_rss
from ..utils import kspace as _k


def padding_lr(enc_x: int, max_pe_index: int) -> Tuple[int, int]:
    """``get_padding`` + the floor/ceil rule of ``zero_pad_kspace_hdr`` from the two header numbers
    (encodedSpace.matrixSize.x, encodingLimits.kspace_encoding_step_1.maximum): 640, 450 -> (94, 95)."""
    p = (enc_x - (max_pe_index + 1)) / 2
    if p % 2 != 0:
        return int(np.floor(p)), int(np.ceil(p))
    return int(p), int(p)


def ifftnd(kspace: Any, axes: Optional[Sequence[int]] = (-1,)) -> Any:
    """Centred inverse FFT along ``axes`` scaled to orthonormal (``* sqrt(prod(shape[axes]))``), signature and
    default ``axes=[-1]`` of ``ZIP!/fastmri_prostate/reconstruction/utils.py:7-29``; ``axes=None`` transforms every
    axis.  The transform is separable: the last two axes go through one 2-D call when both are listed (the path's
    own use, ``prostate_t2_recon.py:99``: ``ifftnd(data_sl, [1, 2])`` on ``(C, RO, PE)``), any other axis is moved
    last and transformed as a batch of lines.  dtype follows the input (complex128 -> complex128)."""
    nd = kspace.ndim
    ax = sorted({a % nd for a in (range(nd) if axes is None else axes)})
    if len(ax) != len(list(range(nd) if axes is None else axes)):
        raise ValueError("repeated axis")
    is_np = isinstance(kspace, np.ndarray)
    xp_move = np.moveaxis if is_np else torch.movedim
    out = kspace
    if nd >= 2 and nd - 2 in ax and nd - 1 in ax:
        out = _k.ifft2c(out)
        ax = [a for a in ax if a < nd - 2]
    for a in ax:
        moved = xp_move(out, a, -1)
        shp = moved.shape
        lines = moved.reshape((-1, 1, shp[-1]))          # (B, 1, N): a length-1 centred transform is the identity
        res = _k.ifft2c(lines).reshape(shp)
        out = xp_move(res, -1, a)
    if is_np:
        return np.ascontiguousarray(out)
    return out.contiguous()


def flip_im(vol: Any, slice_axis: int) -> Any:
    """``ZIP!/fastmri_prostate/reconstruction/utils.py:32-51``, quirk included: ``vol[i] = np.flipud(vol[i])`` for
    ``i < vol.shape[slice_axis]`` indexes the FIRST axis whatever ``slice_axis`` says; in place, returns ``vol``.
    Pure indexing (bit-exact); the fused stage applies the same flip as a store index map (``flip_rows=True``)."""
    for i in range(vol.shape[slice_axis]):
        vol[i] = np.flipud(vol[i]) if isinstance(vol, np.ndarray) else torch.flip(vol[i], dims=(0,))
    return vol


def center_crop_im(im_3d: Any, crop_to_size: Sequence[int]) -> Any:
    """``(slices, y, x)`` centre crop, a view (``ZIP!/fastmri_prostate/reconstruction/utils.py:54-73``):
    ``crop_to_size[0]`` is the x (last axis) size, ``crop_to_size[1]`` the y size; starts truncate ``n/2 - out/2``."""
    x_crop = im_3d.shape[-1] / 2 - crop_to_size[0] / 2
    y_crop = im_3d.shape[-2] / 2 - crop_to_size[1] / 2
    return im_3d[:, int(y_crop):int(crop_to_size[1] + y_crop), int(x_crop):int(crop_to_size[0] + x_crop)]


def rss(sig: Any, axis: int = -1) -> Any:
    """``sqrt(sum(abs(sig)**2, axis))`` for complex (or real) numpy / torch data
    (``ZIP!/fastmri_prostate/reconstruction/t2/prostate_t2_recon.py:105-121``); float32 for complex64 input,
    float64 for complex128 -- numpy's own promotion."""
    is_complex = np.iscomplexobj(sig) if isinstance(sig, np.ndarray) else sig.is_complex()
    mv = D.to_device_complex(sig, name="sig") if is_complex else D.to_device_real(sig, name="sig")
    wide_real = (not is_complex) and (sig.dtype == np.float64 if isinstance(sig, np.ndarray) else sig.dtype == torch.float64)
    out = _cc._rss(mv.tensor, axis, is_complex)
    res = mv.back(out, widen=True)
    if wide_real:
        res = res.astype(np.float64) if isinstance(res, np.ndarray) else res.to(torch.float64)
    return res


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
