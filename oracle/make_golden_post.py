"""Freeze outputs of the UNMODIFIED reference's post-reconstruction steps into ``tests/golden/post_vectors.npz``.

Run in the build container (needs ``/root/reference``)::

    python -m oracle.make_golden_post

Calls ``MRIKneePreprocessor._percentile_clip / _resize_np / _zscore_in_mask / _preview_01``
(``REF/src/preprocess/mri_preprocess.py:182-191,216-233``, imported through ``oracle/ref_shim.py`` with scikit-image stubbed)
in the order ``preprocess_record`` (``:62-84``) calls them, on seeded images (``synth.POST_CASES``) with a stand-in body
mask.  The large cases store every second pixel of the float images (and the full uint8 masks).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mri_acl_imagesegmentation_adsp_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "post_vectors.npz")
CLIP = (1.0, 99.5)            # the reference's default clip_percentiles (mri_preprocess.py:33)


def reference_post(img, mk, out_hw, pre):
    clipped = pre._percentile_clip(img, *CLIP)
    lo, hi = np.percentile(img, CLIP[0]), np.percentile(img, CLIP[1])
    img_r = pre._resize_np(clipped, out_hw)
    mk_r = (pre._resize_np(mk.astype(np.float32), out_hw) > 0.5).astype(np.uint8)
    return {"lo_hi": np.array([lo, hi], dtype=np.float32), "clipped": clipped.astype(np.float32), "img_r": img_r, "mask_r": mk_r,
            "img_z": pre._zscore_in_mask(img_r, mk_r).astype(np.float32), "img_01": pre._preview_01(img_r, mk_r).astype(np.float32)}


def main() -> None:
    if not ref_shim.available():
        raise SystemExit("reference checkout not found; golden vectors can only be made in the build container")
    pre = ref_shim.knee_preprocessor_cls()
    vec, man = {}, {"numpy": np.__version__, "torch": torch.__version__, "clip_percentiles": list(CLIP), "cases": {}}
    for name, shape, out_hw, seed, rule in synth.POST_CASES:
        img, mk, _ = synth.post_case_inputs(name)
        r = reference_post(img, mk, out_hw, pre)
        sub = 2 if shape[0] * shape[1] > 50000 else 1
        vec[f"{name}/lo_hi"] = r["lo_hi"]
        vec[f"{name}/mask_r"] = r["mask_r"]
        for k in ("img_r", "img_z", "img_01"):
            vec[f"{name}/{k}"] = r[k][::sub, ::sub]
        vec[f"{name}/clipped"] = r["clipped"][::4 * sub - 3 if sub > 1 else 1, ::sub]
        man["cases"][name] = {"shape": list(shape), "out": list(out_hw), "seed": seed, "mask": rule, "float_subsample": sub,
                              "clipped_row_step": 4 * sub - 3 if sub > 1 else 1}
    np.savez_compressed(OUT, **vec)
    with open(os.path.join(ROOT, "tests", "golden", "post_manifest.json"), "w") as f:
        json.dump(man, f, indent=1)
    print(f"wrote {OUT}: {len(vec)} arrays, {os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
