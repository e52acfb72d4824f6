"""Cartesian zero-filled reconstruction -- fills the reference's empty ``src/recon/cartesian.py``.

``zero_filled_rss`` is the fused k-space -> image input stage of BASELINE.json's north star:
undersampling-mask apply, (zero-pad of the phase-encode axis), centred 2-D inverse FFT per coil,
root-sum-of-squares coil combine, (flipud, mean over averages), centre crop and instance
normalisation, as ONE C-ABI call per batch (``mriacl_recon_rss_f32``).

It replaces the chain the reference spells three ways (none of them batched or on the GPU):
``src/utils/kspace.py:11-31`` (+ sqrt-sum-squares), the vendored fastMRI functions
``ZIP!/DL_reconstruction/fftc.py:41-65`` / ``coil_combine.py:28-41`` / ``data/transforms.py:45-67,143-162``
and the prostate T2 chain ``ZIP!/fastmri_prostate/reconstruction/t2/prostate_t2_recon.py:65-121``.
The output layout is what ``src/preprocess/mri_preprocess.py:124-140`` and
``src/dataio/datasets.py:90-95`` feed the U-Net: float32, ``(S, oh, ow)`` (``[:, None]`` gives NCHW).
"""
from __future__ import annotations

from typing import Any, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _device as D
from ..adapters import recon_cabi as cabi


def _crop_start(n: int, out: int) -> int:
    return (n - out) // 2


def zero_filled_rss(kspace: Any, mask: Any = None, crop: Optional[Tuple[int, int]] = (320, 320),
                    normalize: Optional[str] = "instance", eps: float = 0.0, flip_rows: bool = False,
                    average_axis: Optional[int] = None, pad: Optional[Tuple[int, int]] = None,
                    *, chunk_slices: Optional[int] = None, force_generic: bool = False, sequential: bool = False,
                    schedule: Optional[str] = None, packed: bool = False, stats_2col: bool = False):
    """k-space -> cropped (normalised) RSS magnitude images.

    kspace   complex64 ``(C,H,W)``, ``(S,C,H,W)`` or, with ``average_axis`` 0 or 1,
             ``(A,S,C,H,W)`` / ``(S,A,C,H,W)``; numpy, torch complex or torch real view ``(...,2)``;
             CPU inputs are copied to the current CUDA device, CUDA inputs are used in place.
    mask     sampling mask along W (0/1 or weights), length W, or None.  (The record dict's
             ``'mask'`` key of the reference is a segmentation mask -- a different thing.)
    crop     ``(oh, ow)`` centre crop, start ``(n-out)//2``; None keeps ``(H, W_padded)``.
             A crop larger than the image raises ValueError like ``center_crop``.
    normalize  ``"instance"`` -> ``(x-mean)/(std+eps)`` with the unbiased std; None -> raw RSS.
    flip_rows  ``np.flipud`` of every combined image (prostate chain).
    average_axis  mean of the per-average RSS images (after the coil combine).
    pad      ``(left, right)`` zero-padding of the W axis before the transform.
    chunk_slices / schedule / force_generic  tuning and testing knobs: slices the workspace holds, kernel
             schedule of the fused plan ("sequential" default, "fused", "overlapped"), generic kernels.
    packed   the last axis of ``kspace`` holds only the sampled columns (``mask != 0``), densely, as written by
             ``mriacl_pack_columns_host`` (``recon.pipeline.HostPipeline`` ships host k-space this way); ``mask`` is
             still the full-width mask.  Shapes with the 640-row column pass only.

    stats_2col  return ``(image, mean_std)`` with the library's own contiguous ``(S, 2)`` statistics tensor instead of two
             strided views of it (``HostPipeline`` copies it to the host in one asynchronous transfer).

    Returns ``(image, mean, std)``: image float32 ``(S,oh,ow)`` (``(oh,ow)`` for a single slice),
    mean/std float32 ``(S,)`` (0-d for a single slice) of the un-normalised crop.  Types follow
    the input (numpy in -> numpy out, CPU torch in -> CPU torch out, CUDA in -> CUDA out).
    """
    if normalize not in (None, "instance"):
        raise ValueError(f"normalize must be None or 'instance', got {normalize!r}")
    mv = D.to_device_complex(kspace)
    k = mv.tensor
    single = False
    if average_axis is None:
        if k.ndim == 3:
            k, single = k[None], True
        if k.ndim != 4:
            raise ValueError(f"kspace must be (C,H,W) or (S,C,H,W), got {tuple(k.shape)}")
        S, C, H, W = k.shape
        A = 1
        slice_stride, avg_stride = C * H * W, 0
    else:
        if k.ndim != 5 or average_axis not in (0, 1):
            raise ValueError("with average_axis the kspace must be (A,S,C,H,W) [axis 0] or (S,A,C,H,W) [axis 1]")
        if average_axis == 0:
            A, S, C, H, W = k.shape
            slice_stride, avg_stride = C * H * W, S * C * H * W
        else:
            S, A, C, H, W = k.shape
            slice_stride, avg_stride = A * C * H * W, C * H * W
    if packed:
        if mask is None:
            raise ValueError("packed k-space needs the sampling mask that selected its columns")
        m_arr = mask.detach().cpu().numpy() if isinstance(mask, torch.Tensor) else np.asarray(mask)
        m_full = D.host_mask(m_arr, int(m_arr.size))
        n_act = int(np.count_nonzero(m_full))
        if W != n_act:
            raise ValueError(f"packed k-space has {W} columns, the mask samples {n_act}")
        # strides are in elements of the packed buffer; W becomes the full line length again
        Wk, W = W, int(m_full.shape[0])
        if average_axis is None:
            slice_stride = C * H * Wk
        elif average_axis == 0:
            slice_stride, avg_stride = C * H * Wk, S * C * H * Wk
        else:
            slice_stride, avg_stride = A * C * H * Wk, C * H * Wk
    pad_left, pad_right = (0, 0) if pad is None else (int(pad[0]), int(pad[1]))
    if pad_left < 0 or pad_right < 0:
        raise ValueError("pad must be non-negative")
    Wp = W + pad_left + pad_right
    oh, ow = (H, Wp) if crop is None else (int(crop[0]), int(crop[1]))
    if not (0 < oh <= H and 0 < ow <= Wp):
        raise ValueError("Invalid shapes.")
    m = D.host_mask(mask, W)
    flags = (cabi.NORM_INSTANCE if normalize == "instance" else 0) | (cabi.FLIP_ROWS if flip_rows else 0) \
        | (cabi.FORCE_GENERIC if force_generic else 0) | (cabi.SEQUENTIAL if sequential else 0) \
        | (cabi.PACKED_COLUMNS if packed else 0)
    if schedule is not None:
        try:
            flags |= {"sequential": cabi.SEQUENTIAL, "fused": cabi.SCHED_FUSED, "overlapped": cabi.SCHED_OVERLAP,
                      "pair": cabi.SCHED_PAIR, "coresident": cabi.SCHED_CORESIDENT,
                      "pipelined": cabi.SCHED_PIPELINED}[schedule]
        except KeyError:
            raise ValueError(f"schedule must be sequential, fused, overlapped, pair, coresident or pipelined, got {schedule!r}") from None

    lib = D.lib()
    out = torch.empty((S, oh, ow), dtype=torch.float32, device=k.device)
    mean_std = torch.empty((S, 2), dtype=torch.float32, device=k.device)
    if S > 0:
        chunk = max(1, min(S, chunk_slices or D.DEFAULT_CHUNK_SLICES))
        nbytes = lib.recon_rss_workspace_bytes(chunk, A, C, H, W, pad_left, Wp, oh, ow, m, flags)
        ws = D.workspace(nbytes)
        lib.recon_rss(k.data_ptr(), slice_stride, avg_stride, m, out.data_ptr(), mean_std.data_ptr(),
                      S, A, C, H, W, pad_left, Wp, oh, ow, flags, float(eps), ws.data_ptr(), ws.numel(), D.stream_ptr())
    if stats_2col:
        return mv.back(out), mv.back(mean_std)
    mean, std = mean_std[:, 0], mean_std[:, 1]
    if single:
        out, mean, std = out[0], mean[0], std[0]
    return mv.back(out), mv.back(mean), mv.back(std)


def recon_to_unet_input(kspace: Any, mask: Any = None, crop: Tuple[int, int] = (320, 320), eps: float = 0.0,
                        **kw) -> Any:
    """``(S,C,H,W)`` k-space -> ``(S,1,oh,ow)`` float32 contiguous NCHW, the tensor contract of
    ``preprocess_records`` (``src/preprocess/mri_preprocess.py:135-140``) / ``KneeNPZ2DSlices``
    (``src/dataio/datasets.py:90-95,133``)."""
    img, _, _ = zero_filled_rss(kspace, mask, crop, "instance", eps, **kw)
    return img[:, None] if isinstance(img, torch.Tensor) else img[:, None, :, :]


#: smp.encoders.get_preprocessing_params("resnet34") with ImageNet weights (REF/src/dataio/datasets.py:69-73)
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def stack_2p5d(volume: Any, k: int = 1, imagenet_norm: bool = False, mean: Optional[Sequence[float]] = None,
               std: Optional[Sequence[float]] = None) -> Any:
    """The network-input epilogue of ``KneeNPZ2DSlices.__getitem__`` (``src/dataio/datasets.py:90-95,128-131``) for a whole
    volume at once: ``(S,1,H,W)`` or ``(S,H,W)`` float32 -> ``(S,k,H,W)`` where channel ``d`` of slice ``s`` is slice
    ``clamp(s + d - k//2, 0, S-1)``; with ``imagenet_norm`` a single channel is repeated to three and every channel is
    normalised ``(x - mean) / std`` (defaults: the ImageNet parameters of smp's resnet encoders).  Pure indexing is bit-exact."""
    mv = D.to_device_real(volume, name="volume")
    x = mv.tensor
    if x.ndim == 4:
        if x.shape[1] != 1:
            raise ValueError(f"volume must be (S,1,H,W) or (S,H,W), got {tuple(x.shape)}")
        x = x[:, 0]
    if x.ndim != 3:
        raise ValueError(f"volume must be (S,1,H,W) or (S,H,W), got {tuple(volume.shape)}")
    k = int(k)
    if k < 1 or k % 2 == 0:
        raise ValueError("k must be a positive odd number of slices")
    repeat = bool(imagenet_norm and k == 1)
    kout = 3 if repeat else k
    S, H, W = x.shape
    mean_t = std_t = None
    if imagenet_norm:
        m = tuple(IMAGENET_MEAN if mean is None else mean)
        sd = tuple(IMAGENET_STD if std is None else std)
        if len(m) != kout or len(sd) != kout:
            raise ValueError(f"mean / std need {kout} entries")          # (torch would fail to broadcast the same way)
        mean_t = torch.tensor(m, dtype=torch.float32, device=x.device)
        std_t = torch.tensor(sd, dtype=torch.float32, device=x.device)
    out = torch.empty((S, kout, H, W), dtype=torch.float32, device=x.device)
    if S:
        D.lib().stack25d(x.contiguous().data_ptr(), out.data_ptr(), S, H * W, kout, repeat, mean_t.data_ptr() if mean_t is not None else 0,
                         std_t.data_ptr() if std_t is not None else 0, D.stream_ptr())
    return mv.back(out)
