"""Twins of the vendored prostate T2 reconstruction tail
(``ZIP!/fastmri_prostate/reconstruction/t2/prostate_t2_recon.py:65-121``,
``reconstruction/utils.py:7-29,54-73``, ``data/mri_data.py:63-85,123-160``) and of the whole ``t2_reconstruction``
(``prostate_t2_recon.py:9-78``): GRAPPA fill per average (``prostate/grappa.py``) -> zero-pad -> coil combine -> mean over
averages -> crop, with the data staying on the device between the GRAPPA launch and the fused stage."""
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


def get_padding(hdr: Any) -> float:
    """``(encodedSpace.matrixSize.x - (kspace_encoding_step_1.maximum + 1)) / 2`` from the ISMRMRD XML header
    (``ZIP!/fastmri_prostate/data/mri_data.py:63-85``; same namespace-qualified descendant query as ``et_query``)."""
    import xml.etree.ElementTree as etree
    root = etree.fromstring(hdr)
    ns = {"n": "http://www.ismrm.org/ISMRMRD"}

    def query(path):
        el = root.find("." + "".join(f"//n:{p}" for p in path), ns)
        if el is None:
            raise RuntimeError("Element not found")
        return str(el.text)

    enc_limits_max = int(query(["encoding", "encodingLimits", "kspace_encoding_step_1", "maximum"])) + 1
    enc_x = int(query(["encoding", "encodedSpace", "matrixSize", "x"]))
    return (enc_x - enc_limits_max) / 2


def _padding_pair(padding: float) -> Tuple[int, int]:
    if padding % 2 != 0:                      # the reference's test (``mri_data.py:152``): floor / ceil unless an even number
        return int(np.floor(padding)), int(np.ceil(padding))
    return int(padding), int(padding)


def zero_pad_kspace_hdr(hdr: Any, unpadded_kspace: Any) -> Any:
    """``np.pad`` of the LAST axis by the header's padding (``mri_data.py:123-160``).  Pure data movement, kept for
    completeness: the fused stage takes ``pad=`` and never materialises the padded array."""
    left, right = _padding_pair(get_padding(hdr))
    if isinstance(unpadded_kspace, torch.Tensor):
        return torch.nn.functional.pad(unpadded_kspace, (left, right))
    return np.pad(unpadded_kspace, ((0, 0),) * (unpadded_kspace.ndim - 1) + ((left, right),))


def t2_reconstruction(kspace_data: Any, calib_data: Any, hdr: Any, crop: Tuple[int, int] = (320, 320)) -> dict:
    """Twin of ``t2_reconstruction`` (``prostate_t2_recon.py:9-78``): ``kspace_data`` complex ``(3, S, C, RO, PE)`` (three
    averages; the second is sampled on the other set of lines), ``calib_data`` ``(S, C, RO, PE_cal)``, ``hdr`` the ISMRMRD XML
    string (or a ``(left, right)`` padding pair).  Returns ``{'reconstruction_rss': (S, 320, 320) float64}``.

    Same steps in the same order: two ``Grappa`` objects from slice 0 of averages 0 and 1; per-slice weights from the slice's
    calibration lines for both (host numpy, ``compute_weights``); averages 0, 1, 2 are filled with objects 1, 2, 1
    (``:52-63``) -- one launch per average on the file's ``(S, C, RO, PE)`` layout; then pad, iFFT, RSS, flipud, mean over
    averages and crop as ONE call of the fused stage on the filled, device-resident k-space."""
    from .grappa import Grappa
    is_np = isinstance(kspace_data, np.ndarray)
    k_host = kspace_data if is_np else kspace_data.detach().cpu().numpy()
    calib = calib_data if isinstance(calib_data, np.ndarray) else calib_data.detach().cpu().numpy()
    if k_host.ndim != 5 or k_host.shape[0] != 3:
        raise ValueError(f"kspace_data must be (3, S, C, RO, PE), got {tuple(k_host.shape)}")
    num_slices = k_host.shape[1]
    g1 = Grappa(np.transpose(k_host[0, 0], (2, 0, 1)), kernel_size=(5, 5), coil_axis=1)
    g2 = Grappa(np.transpose(k_host[1, 0], (2, 0, 1)), kernel_size=(5, 5), coil_axis=1)
    w1 = [g1.compute_weights(np.transpose(calib[s], (2, 0, 1))) for s in range(num_slices)]
    w2 = [g2.compute_weights(np.transpose(calib[s], (2, 0, 1))) for s in range(num_slices)]
    dev = D.require_cuda()
    k_dev = torch.from_numpy(np.ascontiguousarray(k_host, dtype=np.complex64)).to(dev) if is_np else \
        kspace_data.to(device=dev, dtype=torch.complex64).contiguous()
    filled = torch.empty_like(k_dev)
    for average, (g, w) in enumerate(((g1, w1), (g2, w2), (g1, w1))):
        filled[average] = g.apply_weights_batch(k_dev[average], w, axes=(2, 1, 0))        # x = PE, y = RO, coil = axis 0
    pad = _padding_pair(get_padding(hdr)) if isinstance(hdr, (str, bytes)) else (int(hdr[0]), int(hdr[1]))
    img = t2_average_combine(filled, pad, crop, None)
    img = img.cpu().numpy() if is_np else (img if kspace_data.device.type == "cuda" else img.cpu())
    return {"reconstruction_rss": img}
