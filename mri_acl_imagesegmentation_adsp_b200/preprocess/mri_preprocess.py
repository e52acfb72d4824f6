"""The k-space branch of the reference's ``MRIKneePreprocessor``
(``src/preprocess/mri_preprocess.py``), on the GPU.

Only the reconstruction step is in scope: ``ifft2c_single`` -- the one
variant on the reference's live call path (``:59``) -- and batched forms that keep the
``(S,1,H,W)`` float32 tensor contract of ``preprocess_records`` (``:124-140``).  Percentile clip, Otsu
body mask, N4, NL-means, resize and in-mask z-score stay in the reference's Python.
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

    def __init__(self, out_size: Tuple[int, int] = (320, 320)) -> None:
        self.out_size = out_size

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
