"""The data format on the OUTPUT side of the path: the per-volume artefacts the training code
reads back (``REF/src/dataio/datasets.py`` loads ``volume.npz``; ``REF/src/train/train_unet.py`` globs them).

Twins of ``REF/src/main.py:85-149`` (``group_records_by_file``, ``save_pack``) and of the volume loop of ``build_preprocess``
(``:183-214``), fed by the device preprocessor (``preprocess.mri_preprocess.MRIKneePreprocessor.preprocess_records``):
``tensor.pt`` (S,1,H,W) float32, ``volume.npz`` {img float32 (S,1,H,W), msk uint8 (S,H,W)}, ``mask.npy``, ``indices.json``,
``metas.json``, ``preview/slice_XXX.png`` (8-bit greyscale; written with Pillow -- the reference uses imageio, same pixels) and
``stats.json`` (in-mask mean / population std per slice, first 50)."""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np
import torch


def group_records_by_file(records: List[Dict]) -> Dict[str, List[Dict]]:
    buckets: Dict[str, List[Dict]] = {}
    for r in records:
        buckets.setdefault(r["filepath"], []).append(r)
    for fp in buckets:
        buckets[fp] = sorted(buckets[fp], key=lambda x: x["slice_idx"])
    return buckets


def _write_png(path: str, img_u8: np.ndarray) -> None:
    from PIL import Image
    Image.fromarray(img_u8).save(path)


def save_pack(out_dir: str, pack: Dict[str, Any], preview_max: int = 8) -> None:
    os.makedirs(out_dir, exist_ok=True)
    tensor: torch.Tensor = pack["tensor"]
    torch.save(tensor, os.path.join(out_dir, "tensor.pt"))
    volume_np = tensor.detach().cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "volume.npz"), img=volume_np.astype(np.float32, copy=False),
                        msk=pack["mask"].astype(np.uint8, copy=False))
    np.save(os.path.join(out_dir, "mask.npy"), pack["mask"])
    with open(os.path.join(out_dir, "indices.json"), "w", encoding="utf-8") as f:
        json.dump(pack.get("indices", []), f, ensure_ascii=False, indent=2)
    with open(os.path.join(out_dir, "metas.json"), "w", encoding="utf-8") as f:
        json.dump(pack.get("metas", []), f, ensure_ascii=False, indent=2)
    prev = pack["preview"]
    pv_dir = os.path.join(out_dir, "preview")
    os.makedirs(pv_dir, exist_ok=True)
    S = prev.shape[0]
    for i in range(min(preview_max, S)):
        _write_png(os.path.join(pv_dir, f"slice_{pack['indices'][i]:03d}.png"), (prev[i] * 255).astype(np.uint8))
    img_z = tensor[:, 0].numpy()
    mk = pack["mask"]
    means, stds = [], []
    for s in range(img_z.shape[0]):
        vals = img_z[s][mk[s] > 0]
        if vals.size == 0:
            means.append(float("nan")); stds.append(float("nan"))
        else:
            means.append(float(vals.mean())); stds.append(float(vals.std()))
    stats = {"count_slices": int(S), "mean_in_mask_mean": float(np.nanmean(means)), "mean_in_mask_std": float(np.nanmean(stds)),
             "per_slice_mean": means[:50], "per_slice_std": stds[:50]}
    with open(os.path.join(out_dir, "stats.json"), "w", encoding="utf-8") as f:
        json.dump(stats, f, ensure_ascii=False, indent=2)


def preprocess_volumes(adapter: Any, out_dir: str, preprocessor: Any, root_dir: Optional[str] = None, preview_max: int = 8) -> List[Dict]:
    """The volume loop of ``build_preprocess`` (``REF/src/main.py:196-214``): discover -> group by file -> load -> ONE device
    call per volume (``preprocessor.preprocess_records``) -> ``save_pack``.  Returns the same summary list."""
    out_root = Path(out_dir)
    out_root.mkdir(parents=True, exist_ok=True)
    try:
        records = adapter.discover_records(root_dir)
    except TypeError:
        records = adapter.discover_records()
    summary = []
    for filepath, record_defs in group_records_by_file(records).items():
        loaded = [adapter.load_record(rec) for rec in record_defs]
        pack = preprocessor.preprocess_records(loaded)
        vol_dir = out_root / Path(filepath).stem
        save_pack(str(vol_dir), pack, preview_max=preview_max)
        summary.append({"filepath": filepath, "output_dir": str(vol_dir), "npz_path": str(vol_dir / "volume.npz"),
                        "num_slices": int(pack["tensor"].shape[0])})
    return summary
