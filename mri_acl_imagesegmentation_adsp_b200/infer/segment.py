"""Inference driver for BASELINE.json configs[3]: multicoil k-space -> fused input stage -> U-Net -> mask.

Fills the empty ``REF/src/infer/segment.py`` (``REF/src/guide.txt:86-87`` refers to it).  The network input
is exactly the tensor contract of ``preprocess_records`` / ``KneeNPZ2DSlices``
(``REF/src/preprocess/mri_preprocess.py:135-140``, ``REF/src/dataio/datasets.py:90-95,133``): float32
``(B,1,320,320)``, produced on the device by the fused stage and never copied to the host; the prediction
rule is ``sigmoid(logits) > 0.5`` as in ``REF/src/train/engine.py:132``.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import torch

from ..recon.cartesian import recon_to_unet_input


@torch.no_grad()
def segment_kspace(model: torch.nn.Module, kspace: Any, sampling_mask: Any = None, crop: Tuple[int, int] = (320, 320),
                   threshold: float = 0.5, amp: bool = True, net_batch: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """``kspace``: CUDA complex64 ``(S,C,H,W)``.  Returns ``{"input": (S,1,oh,ow) f32, "logits": (S,K,oh,ow) f32,
    "mask": (S,K,oh,ow) bool}``, all on the device."""
    x = recon_to_unet_input(kspace, sampling_mask, crop)
    if not isinstance(x, torch.Tensor) or x.device.type != "cuda":
        raise ValueError("segment_kspace expects device-resident k-space (torch CUDA tensor)")
    model.eval()
    step = net_batch or x.shape[0]
    outs = []
    for s0 in range(0, x.shape[0], step):
        with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            outs.append(model(x[s0:s0 + step]).float())
    logits = torch.cat(outs, 0) if len(outs) > 1 else outs[0]
    return {"input": x, "logits": logits, "mask": torch.sigmoid(logits) > threshold}
