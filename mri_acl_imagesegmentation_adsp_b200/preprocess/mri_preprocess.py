"""The k-space branch of the reference's ``MRIKneePreprocessor``
(``src/preprocess/mri_preprocess.py``), on the GPU.

The reconstruction step: ``ifft2c_single`` -- the one variant on the reference's
live call path (``:59``) -- and batched forms that keep the ``(S,1,H,W)`` float32 tensor contract of
``preprocess_records`` (``:124-140``); and the per-slice steps that follow it on that path (SURVEY.md section 8f
row 2): percentile clip (``:182-185``), bilinear resize (``:187-191``), in-mask z-score (``:216-224``) and the
``[0,1]`` preview (``:226-233``), each as a twin of the reference's static method and together as one device call
per volume (``preprocess_records``).  The Otsu / morphology body mask (``:194-214``) needs scikit-image, N4 needs
SimpleITK and NL-means scikit-image again: those stay in the reference's Python -- the body mask comes in through
the record (``'body_mask'``) or a callable, and without one every pixel counts as inside.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _device as D
from ..recon.cartesian import zero_filled_rss


def _ifft2c_abs_batch(k: torch.Tensor) -> torch.Tensor:
    b, h, w = k.shape
    out = torch.empty((b, h, w), dtype=torch.float32, device=k.device)
    if b:
        lib = D.lib()
        nbytes = lib.ifft2c_abs_workspace_bytes(b, h, w)
        ws = D.workspace(nbytes)
        lib.ifft2c_abs(k.data_ptr(), out.data_ptr(), b, h, w, ws.data_ptr(), ws.numel(), D.stream_ptr())
    return out


class MRIKneePreprocessor:
    """Reconstruction part of the reference class of the same name."""

    def __init__(self, out_size: Tuple[int, int] = (320, 320), slice_keep: Tuple[float, float] = (0.3, 0.7),
                 clip_percentiles: Tuple[float, float] = (1.0, 99.5), use_n4: bool = False, use_denoise: bool = False,
                 body_mask_fn: Optional[Any] = None) -> None:
        """Arguments of the reference constructor (``:28-41``) plus ``body_mask_fn``: a callable ``(H,W) float32 numpy ->
        (H,W) uint8`` applied to each CLIPPED full-resolution image on the host (hand it the reference's ``_body_mask``
        where scikit-image is installed).  ``use_n4`` / ``use_denoise`` are not available on the device."""
        self.out_size = tuple(out_size)
        self.slice_keep = tuple(slice_keep)
        self.clip_percentiles = tuple(clip_percentiles)
        if use_n4 or use_denoise:
            raise ValueError("N4 bias correction and NL-means denoising are not part of the device path")
        self.body_mask_fn = body_mask_fn
        lo, hi = self.slice_keep
        if not (0.0 <= lo < hi <= 1.0):
            raise ValueError("slice_keep must satisfy 0.0 <= lo < hi <= 1.0")                    # mri_preprocess.py:163-166
        pmin, pmax = self.clip_percentiles
        if not (0.0 <= pmin < pmax <= 100.0):
            raise ValueError("clip_percentiles must lie in [0,100] with pmin < pmax")            # :167-169

    @staticmethod
    def _ensure_2d(x: Any, name: str) -> Any:
        if x.ndim != 2:
            raise ValueError(f"{name} must have shape (H,W), got {tuple(x.shape)}")   # mri_preprocess.py:177-180
        return x

    @staticmethod
    def ifft2c_single(kspace_2d: Any) -> Any:
        """Centred 2-D iFFT magnitude of ONE single-coil slice: complex ``(H,W)`` -> float32 ``(H,W)``
        (``src/preprocess/mri_preprocess.py:149-160``).  ValueError when ``ndim != 2``."""
        MRIKneePreprocessor._ensure_2d(kspace_2d, "kspace")
        mv = D.to_device_complex(kspace_2d)
        return mv.back(_ifft2c_abs_batch(mv.tensor[None])[0])

    @staticmethod
    def ifft2c_batch(kspace: Any) -> Any:
        """``(S,H,W)`` complex -> ``(S,H,W)`` float32: ``ifft2c_single`` for a whole volume in one call."""
        mv = D.to_device_complex(kspace)
        if mv.tensor.ndim != 3:
            raise ValueError(f"kspace must have shape (S,H,W), got {tuple(mv.tensor.shape)}")
        return mv.back(_ifft2c_abs_batch(mv.tensor))

    # ---- the steps after the reconstruction: twins of the reference's static methods --------------------------
    @staticmethod
    def _as_batch(x: Any, name: str):
        mv = D.to_device_real(x, name=name)
        t = mv.tensor
        if t.ndim not in (2, 3):
            raise ValueError(f"{name} must have shape (H,W) or (B,H,W), got {tuple(t.shape)}")
        return mv, (t[None] if t.ndim == 2 else t), t.ndim == 2

    @staticmethod
    def _mask_u8(mask: Any, shape: Tuple[int, ...]) -> torch.Tensor:
        dev = D.require_cuda()
        m = torch.from_numpy(np.ascontiguousarray(mask)) if isinstance(mask, np.ndarray) else mask
        m = (m > 0).to(device=dev, dtype=torch.uint8).contiguous()
        if tuple(m.shape) != tuple(shape):
            raise ValueError(f"mask shape {tuple(m.shape)} != image shape {tuple(shape)}")
        return m

    @staticmethod
    def _percentile_clip(img: Any, pmin: float, pmax: float) -> Any:
        """``np.clip(img, np.percentile(img, pmin), np.percentile(img, pmax))`` per image (``:182-185``); exact order
        statistics, numpy's float32 interpolation."""
        mv, t, single = MRIKneePreprocessor._as_batch(img, "img")
        out = torch.empty_like(t)
        b, h, w = t.shape
        D.lib().percentile_clip(t.data_ptr(), out.data_ptr(), 0, b, h * w, float(pmin), float(pmax), D.stream_ptr())
        return mv.back(out[0] if single else out)

    @staticmethod
    def _resize_np(img: Any, out_hw: Tuple[int, int]) -> Any:
        """``F.interpolate(size=out_hw, mode='bilinear', align_corners=False)`` (``:187-191``)."""
        mv, t, single = MRIKneePreprocessor._as_batch(img, "img")
        b, h, w = t.shape
        out = torch.empty((b, int(out_hw[0]), int(out_hw[1])), dtype=torch.float32, device=t.device)
        D.lib().resize_bilinear(t.data_ptr(), out.data_ptr(), b, h, w, int(out_hw[0]), int(out_hw[1]), D.stream_ptr())
        return mv.back(out[0] if single else out)

    @staticmethod
    def _zscore_in_mask(img: Any, mask: Any) -> Any:
        """``(img - mean) / std`` with the statistics of the pixels inside ``mask`` (``:216-224``)."""
        mv, t, single = MRIKneePreprocessor._as_batch(img, "img")
        m = MRIKneePreprocessor._mask_u8(mask, img.shape)
        out = torch.empty_like(t)
        b, h, w = t.shape
        D.lib().zscore_preview(t.data_ptr(), m.data_ptr(), out.data_ptr(), 0, 0, b, h * w, D.stream_ptr())
        return mv.back(out[0] if single else out)

    @staticmethod
    def _preview_01(img: Any, mask: Any) -> Any:
        """``(img - lo) / (hi - lo + 1e-6)`` with the extrema inside ``mask`` (``:226-233``)."""
        mv, t, single = MRIKneePreprocessor._as_batch(img, "img")
        m = MRIKneePreprocessor._mask_u8(mask, img.shape)
        out = torch.empty_like(t)
        b, h, w = t.shape
        D.lib().zscore_preview(t.data_ptr(), m.data_ptr(), 0, out.data_ptr(), 0, b, h * w, D.stream_ptr())
        return mv.back(out[0] if single else out)

    def clip_resize_zscore(self, imgs: torch.Tensor, body_masks: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Device tensors in, device tensors out: ``imgs`` float32 ``(B,H,W)`` (e.g. the output of the fused stage or of
        ``ifft2c_batch``), ``body_masks`` uint8 ``(B,H,W)`` or None -> ``{'img_z', 'img_01': (B,oh,ow) f32, 'mask':
        (B,oh,ow) u8, 'clip': (B,2), 'stats': (B,6)}`` -- ONE C-ABI call (``mriacl_clip_resize_zscore_f32``)."""
        if imgs.device.type != "cuda" or imgs.dtype != torch.float32 or imgs.ndim != 3:
            raise ValueError("imgs must be a CUDA float32 (B,H,W) tensor")
        imgs = imgs.contiguous()
        b, h, w = imgs.shape
        oh, ow = self.out_size
        dev = imgs.device
        z = torch.empty((b, oh, ow), dtype=torch.float32, device=dev)
        p01 = torch.empty_like(z)
        clip = torch.empty((b, 2), dtype=torch.float32, device=dev)
        stats = torch.empty((b, 6), dtype=torch.float32, device=dev)
        if body_masks is not None:
            body_masks = self._mask_u8(body_masks, imgs.shape)
            mk = torch.empty((b, oh, ow), dtype=torch.uint8, device=dev)
        else:
            mk = None
        D.lib().clip_resize_zscore(imgs.data_ptr(), body_masks.data_ptr() if body_masks is not None else 0, z.data_ptr(),
                                   p01.data_ptr(), mk.data_ptr() if mk is not None else 0, clip.data_ptr(), stats.data_ptr(),
                                   b, h, w, oh, ow, float(self.clip_percentiles[0]), float(self.clip_percentiles[1]), D.stream_ptr())
        if mk is None:
            mk = torch.ones((b, oh, ow), dtype=torch.uint8, device=dev)
        return {"img_z": z, "img_01": p01, "mask": mk, "clip": clip, "stats": stats}

    def preprocess_records(self, records: Sequence[Dict[str, Any]]) -> Dict[str, Any]:
        """``preprocess_records`` (``:96-146``) for k-space records on the device: the ``slice_keep`` band, reconstruction,
        clip, (body mask: ``record['body_mask']`` at full resolution, else ``body_mask_fn`` on the host, else all inside),
        resize, z-score, preview.  Same output dict: ``'tensor'`` (S,1,H,W) float32 torch (CPU, as the reference returns
        it), ``'preview'`` / ``'mask'`` numpy, ``'indices'``, ``'sources'``, ``'metas'``."""
        ns = len(records)
        if ns == 0:
            raise ValueError("No records provided to preprocess_records.")
        s0 = max(0, int(ns * self.slice_keep[0]))
        s1 = min(ns, int(ns * self.slice_keep[1]))
        s1 = max(s1, s0 + 1)
        if s1 > ns:
            s1 = ns
        if s0 >= s1:
            s0, s1 = 0, ns
        recs = list(records[s0:s1])
        # source precedence of _normalize_record_input (:262-296): image -> target / reconstruction* -> kspace
        dev = D.require_cuda()
        sources, planes, k_idx, ks = [], [None] * len(recs), [], []
        for i, r in enumerate(recs):
            if r.get("image", None) is not None:
                planes[i] = self._ensure_2d(np.squeeze(np.asarray(r["image"])).astype(np.float32, copy=False), "image")
                sources.append("image")
                continue
            for key in ("target", "reconstruction", "reconstruction_rss", "reconstruction_esc"):
                if r.get(key, None) is not None:
                    planes[i] = self._ensure_2d(np.squeeze(np.asarray(r[key])).astype(np.float32, copy=False), key)
                    sources.append("target")
                    break
            else:
                ks.append(self._record_kspace(r))
                k_idx.append(i)
                sources.append("kspace")
        if ks:
            shapes = {tuple(k.shape) for k in ks}
            if len(shapes) != 1:
                raise ValueError(f"records of one volume must share a shape, got {sorted(shapes)}")
            vol = np.stack([np.asarray(k, dtype=np.complex64) for k in ks]) if isinstance(ks[0], np.ndarray) else torch.stack(list(ks))
            rec_imgs = _ifft2c_abs_batch(D.to_device_complex(vol).tensor)       # ONE device call for all k-space records
        if len(ks) == len(recs):
            imgs = rec_imgs
        else:
            shapes = {tuple(p.shape) for p in planes if p is not None} | ({tuple(rec_imgs.shape[1:])} if ks else set())
            if len(shapes) != 1:
                raise ValueError(f"records of one volume must share a shape, got {sorted(shapes)}")
            imgs = torch.empty((len(recs),) + next(iter(shapes)), dtype=torch.float32, device=dev)
            for j, i in enumerate(k_idx):
                imgs[i] = rec_imgs[j]
            for i, p in enumerate(planes):
                if p is not None:
                    imgs[i] = torch.from_numpy(np.ascontiguousarray(p)).to(dev)
        masks = None
        if all(r.get("body_mask", None) is not None for r in recs):
            masks = torch.from_numpy(np.stack([np.asarray(r["body_mask"]) for r in recs]))
        elif self.body_mask_fn is not None:
            clipped = self._percentile_clip(imgs, *self.clip_percentiles).cpu().numpy()
            masks = torch.from_numpy(np.stack([np.asarray(self.body_mask_fn(c)).astype(np.uint8) for c in clipped]))
        out = self.clip_resize_zscore(imgs, masks)
        metas = [r.get("meta", {}) for r in recs]
        return {"tensor": out["img_z"][:, None].cpu(), "preview": out["img_01"].cpu().numpy(),
                "mask": out["mask"].cpu().numpy(), "indices": [m.get("slice_idx", s0 + i) for i, m in enumerate(metas)],
                "sources": sources, "metas": metas}

    @staticmethod
    def _record_kspace(record: Dict[str, Any]) -> np.ndarray:
        """The k-space branch of ``_normalize_record_input`` (``:285-296``): squeeze, reject a
        real/imag ``(2,H,W)`` stack, require ``(H,W)``."""
        ksp = record.get("kspace", None)
        if ksp is None:
            raise ValueError("record has no 'kspace'")
        ksp = np.squeeze(ksp) if isinstance(ksp, np.ndarray) else ksp.squeeze()
        is_complex = np.iscomplexobj(ksp) if isinstance(ksp, np.ndarray) else ksp.is_complex()
        if not is_complex and ksp.ndim == 3 and ksp.shape[0] == 2:
            raise ValueError("kspace is not complex: combine (real, imag) -> complex before preprocessing")
        return MRIKneePreprocessor._ensure_2d(ksp, "kspace")

    def recon_records(self, records: Sequence[Dict[str, Any]]) -> Dict[str, Any]:
        """Single-coil records ``{'kspace': (H,W) complex, 'meta': ...}`` -> ``{'tensor': (S,1,H,W) f32,
        'metas': [...], 'sources': [...]}`` -- the stacking contract of ``preprocess_records``
        (``:122-146``) for the reconstruction step, one device call for the whole volume."""
        ks = [self._record_kspace(r) for r in records]
        if not ks:
            raise ValueError("no records")
        shapes = {tuple(k.shape) for k in ks}
        if len(shapes) != 1:
            raise ValueError(f"records of one volume must share a shape, got {sorted(shapes)}")
        if isinstance(ks[0], np.ndarray):
            vol = np.stack([np.asarray(k, dtype=np.complex64) for k in ks])
        else:
            vol = torch.stack(list(ks))
        img = self.ifft2c_batch(vol)
        tensor = img[:, None] if isinstance(img, torch.Tensor) else torch.from_numpy(img[:, None])
        return {"tensor": tensor, "metas": [r.get("meta", {}) for r in records],
                "sources": ["kspace"] * len(records), "indices": list(range(len(records)))}


def ifft2c_single(kspace_2d: Any) -> Any:
    return MRIKneePreprocessor.ifft2c_single(kspace_2d)


__all__ = ["MRIKneePreprocessor", "ifft2c_single", "zero_filled_rss"]
