"""Freeze outputs of the UNMODIFIED reference functions into ``tests/golden/``.

Run in the build container (needs ``/root/reference``)::

    python -m oracle.make_golden

The reference has no golden vectors of its own, so these
files ARE the pin: every array below is produced by calling the reference's own
code (``oracle/ref_shim.py``) on seeded synthetic k-space
(``mri_acl_imagesegmentation_adsp_b200/synth.py``).  ``tests/test_oracle_golden.py``
then checks ``oracle/recon_oracle.py`` against them on any machine, and the
``-m gpu`` tests check the CUDA path against both.  Large inputs are not stored:
they are regenerated from (generator, shape, seed) recorded in the manifest.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mri_acl_imagesegmentation_adsp_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT_DIR = os.path.join(ROOT, "tests", "golden")


def _sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main() -> None:
    if not ref_shim.available():
        raise SystemExit("reference checkout not found; golden vectors can only be made in the build container")
    ks = ref_shim.kspace_utils()
    pre = ref_shim.knee_preprocessor_cls()
    dl = ref_shim.fastmri_dl()
    pr = ref_shim.prostate()

    vec: dict[str, np.ndarray] = {}
    man: dict = {"numpy": np.__version__, "torch": torch.__version__, "cases": {}}

    def ref_numpy_chain(k, mask, crop):
        km = k if mask is None else k * mask.reshape((1,) * (k.ndim - 1) + (-1,))
        mag = ks.complex_abs(ks.ifft2c(km))
        return ks.center_crop_or_pad(np.sqrt((mag ** 2).sum(axis=-3)), crop[0], crop[1])

    def ref_fastmri_chain(k, mask, crop):
        t = dl.transforms.to_tensor(k)
        if mask is not None:
            t = t * torch.from_numpy(mask).reshape(1, 1, -1, 1)
        img = dl.coil_combine.rss_complex(dl.fftc.ifft2c_new(t), dim=0)
        img = dl.transforms.center_crop(img, crop)
        out, mean, std = dl.transforms.normalize_instance(img, eps=0.0)
        return out.numpy(), np.float32(mean.item()), np.float32(std.item())

    # ---- 1. small even knee case, inputs stored -------------------------------------
    k = synth.gaussian_kspace((4, 32, 24), seed=11)
    m = synth.equispaced_mask(24, 4, 0.25)
    vec["small_even/kspace"] = k
    vec["small_even/mask"] = m
    vec["small_even/ifft2c"] = ks.ifft2c(k)
    vec["small_even/fft2c"] = ks.fft2c(k)
    vec["small_even/numpy_chain_16x16"] = ref_numpy_chain(k, m, (16, 16))
    o, mu, sd = ref_fastmri_chain(k, m, (16, 16))
    vec["small_even/fastmri_chain_16x16"] = o
    vec["small_even/fastmri_mean_std"] = np.array([mu, sd], dtype=np.float32)
    vec["small_even/ifft2c_single_coil0"] = pre.ifft2c_single(k[0])
    vec["small_even/crop_or_pad_40x12"] = ks.center_crop_or_pad(vec["small_even/ifft2c_single_coil0"], 40, 12)
    vec["small_even/complex_abs"] = ks.complex_abs(k)

    # ---- 2. odd sizes (ifftshift != fftshift) ---------------------------------------
    k = synth.gaussian_kspace((2, 3, 30, 23), seed=12)
    vec["small_odd/kspace"] = k
    vec["small_odd/ifft2c"] = ks.ifft2c(k)
    vec["small_odd/fft2c"] = ks.fft2c(k)
    t = dl.transforms.to_tensor(k)
    vec["small_odd/ifft2c_new"] = dl.fftc.ifft2c_new(t).numpy()
    vec["small_odd/fft2c_new"] = dl.fftc.fft2c_new(t).numpy()
    vec["small_odd/rss_complex"] = dl.coil_combine.rss_complex(dl.fftc.ifft2c_new(t), dim=1).numpy()
    vec["small_odd/rss_real"] = dl.coil_combine.rss(torch.from_numpy(np.abs(k)), dim=1).numpy()
    vec["small_odd/ifft2c_single"] = pre.ifft2c_single(k[1, 2])
    vec["small_odd/ifftnd"] = pr.utils.ifftnd(k[0].copy(), [1, 2])

    # ---- 3. configs[0]: 15-coil 640x368 knee slice, 4x mask, crop 320x320 -----------
    m = synth.knee_mask()
    vec["knee/mask_368_4x_008_idx"] = np.flatnonzero(m).astype(np.int32)
    for name, gen, seed in (("knee_gauss", synth.gaussian_kspace, 0), ("knee_phantom", synth.phantom_kspace, 0)):
        k = gen(synth.KNEE_SHAPE, seed)
        a = ref_numpy_chain(k, m, synth.CROP)
        o, mu, sd = ref_fastmri_chain(k, m, synth.CROP)
        vec[f"{name}/numpy_chain"] = a.astype(np.float32)
        vec[f"{name}/fastmri_chain"] = o
        vec[f"{name}/fastmri_mean_std"] = np.array([mu, sd], dtype=np.float32)
        man["cases"][name] = {"generator": gen.__name__, "shape": list(synth.KNEE_SHAPE), "seed": seed,
                              "mask": "equispaced(368,4,0.08)", "crop": list(synth.CROP),
                              "kspace_sha256": _sha(k), "numpy_chain_sha256": _sha(a)}
    # unmasked (dense) variant of the gaussian slice, numpy chain only, subsampled
    k = synth.gaussian_kspace(synth.KNEE_SHAPE, 0)
    vec["knee_gauss/numpy_chain_nomask_sub4"] = ref_numpy_chain(k, None, synth.CROP)[::4, ::4].astype(np.float32)

    # ---- 4. single-coil live path: ifft2c_single on (640, 368) ----------------------
    k1 = synth.gaussian_kspace((640, 368), seed=1)
    a = pre.ifft2c_single(k1)
    vec["single_coil/ifft2c_single_rows_even"] = a[::2]
    man["cases"]["single_coil"] = {"generator": "gaussian_kspace", "shape": [640, 368], "seed": 1,
                                   "sha256": _sha(a), "stored": "rows ::2"}
    a372 = pre.ifft2c_single(synth.gaussian_kspace((640, 372), seed=2))
    vec["single_coil/ifft2c_single_372_sub"] = a372[::5, ::3]
    man["cases"]["single_coil_372"] = {"generator": "gaussian_kspace", "shape": [640, 372], "seed": 2,
                                       "sha256": _sha(a372), "stored": "[::5, ::3]"}

    # ---- 5. prostate T2 chain (GRAPPA excluded) -------------------------------------
    hdr = ref_shim.synthetic_header(32, 20)
    p = pr.mri_data.get_padding(hdr)
    vec["prostate_small/get_padding"] = np.array([p], dtype=np.float64)
    k = synth.gaussian_kspace((2, 2, 3, 32, 21), seed=13)
    vec["prostate_small/kspace"] = k
    ims = np.zeros((2, 2, 32, 32))
    for av in range(2):
        padded = pr.mri_data.zero_pad_kspace_hdr(hdr, k[av])
        vec[f"prostate_small/padded_shape_{av}"] = np.array(padded.shape, dtype=np.int64)
        ims[av] = pr.t2.create_coil_combined_im(padded)
    vec["prostate_small/coil_combined"] = ims
    vec["prostate_small/final_16x16"] = pr.utils.center_crop_im(np.mean(ims, axis=0), [16, 16])

    hdr = ref_shim.synthetic_header(640, 450)
    vec["prostate/get_padding_640_451"] = np.array([pr.mri_data.get_padding(hdr)], dtype=np.float64)
    pm = synth.prostate_mask()
    vec["prostate/mask_451_8x_004_idx"] = np.flatnonzero(pm).astype(np.int32)
    shape = (3, 1, 16, 640, 451)
    k = synth.gaussian_kspace(shape, seed=0) * pm.reshape(1, 1, 1, 1, -1)
    ims = np.zeros((3, 1, 640, 640))
    for av in range(3):
        ims[av] = pr.t2.create_coil_combined_im(pr.mri_data.zero_pad_kspace_hdr(hdr, k[av]))
    fin = pr.utils.center_crop_im(np.mean(ims, axis=0), [320, 320])
    vec["prostate/one_slice_final"] = fin
    man["cases"]["prostate_one_slice"] = {"generator": "gaussian_kspace", "shape": list(shape), "seed": 0,
                                          "mask": "equispaced(451,8,0.04)", "pad": [94, 95], "sha256": _sha(fin)}

    # ---- 5b. configs[2] at full size: two slices of the (3,30,16,640,451) volume, built block by block ----
    for s_idx in (0, 29):
        kk = np.stack([synth.prostate_volume_block(a, s_idx) for a in range(3)])[:, None] * pm.reshape(1, 1, 1, 1, -1)
        ims = np.zeros((3, 1, 640, 640))
        for av in range(3):
            ims[av] = pr.t2.create_coil_combined_im(pr.mri_data.zero_pad_kspace_hdr(hdr, kk[av]))
        fin = pr.utils.center_crop_im(np.mean(ims, axis=0), [320, 320])
        vec[f"prostate/volume_slice{s_idx}_sub2"] = fin[0, ::2, ::2].astype(np.float32)
        man["cases"][f"prostate_volume_slice{s_idx}"] = {"generator": "prostate_volume_block", "slice": s_idx, "seed": 0,
                                                        "mask": "equispaced(451,8,0.04)", "pad": [94, 95],
                                                        "sha256": _sha(fin), "stored": "[0, ::2, ::2] as float32"}

    # ---- 5c. configs[0] slice under the 8x knee mask and under 4x masks with a non-zero offset ----
    k = synth.gaussian_kspace(synth.KNEE_SHAPE, 0)
    for tag, mm in (("8x", synth.equispaced_mask(368, 8, 0.04)), ("4x_off1", synth.equispaced_mask(368, 4, 0.08, 1)),
                    ("4x_off3", synth.equispaced_mask(368, 4, 0.08, 3))):
        a = ref_numpy_chain(k, mm, synth.CROP)
        o, mu, sd = ref_fastmri_chain(k, mm, synth.CROP)
        vec[f"knee_gauss/numpy_chain_{tag}_sub2"] = a[::2, ::2].astype(np.float32)
        vec[f"knee_gauss/fastmri_mean_std_{tag}"] = np.array([mu, sd], dtype=np.float32)
        man["cases"][f"knee_gauss_{tag}"] = {"generator": "gaussian_kspace", "shape": list(synth.KNEE_SHAPE), "seed": 0,
                                             "columns_kept": int(np.count_nonzero(mm)), "numpy_chain_sha256": _sha(a),
                                             "stored": "[::2, ::2]"}

    # ---- 5d. frozen sampling-mask index lists (the generator is builder-defined: SURVEY.md section 8c) ----
    man["masks"] = {synth.mask_name(*spec): [int(i) for i in np.flatnonzero(synth.equispaced_mask(*spec))]
                    for spec in synth.FROZEN_MASKS}

    # ---- 6. normalize_instance on its own -------------------------------------------
    x = np.abs(synth.gaussian_kspace((320, 320), seed=3)).astype(np.float32) + 2.0
    o, mu, sd = dl.transforms.normalize_instance(torch.from_numpy(x), eps=1e-11)
    vec["norm/out_sub"] = o.numpy()[::8, ::8]
    vec["norm/mean_std"] = np.array([mu.item(), sd.item()], dtype=np.float32)

    # ---- 7. index tables --------------------------------------------------------------
    vec["index/crop_rows_640_320"] = np.array([160, 480], dtype=np.int32)
    vec["index/crop_cols_368_320"] = np.array([24, 344], dtype=np.int32)
    probe = np.arange(640 * 368, dtype=np.float32).reshape(640, 368)
    vec["index/crop_probe_corners"] = ks.center_crop_or_pad(probe, 320, 320)[[0, 0, -1, -1], [0, -1, 0, -1]]
    tprobe = dl.transforms.center_crop(torch.from_numpy(probe), (320, 320)).numpy()
    assert np.array_equal(tprobe, ks.center_crop_or_pad(probe, 320, 320))

    os.makedirs(OUT_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(OUT_DIR, "reference_vectors.npz"), **vec)
    man["arrays"] = {k_: {"shape": list(v.shape), "dtype": str(v.dtype)} for k_, v in vec.items()}
    with open(os.path.join(OUT_DIR, "manifest.json"), "w") as f:
        json.dump(man, f, indent=1, sort_keys=True)
    size = os.path.getsize(os.path.join(OUT_DIR, "reference_vectors.npz"))
    print(f"wrote {len(vec)} arrays, {size / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
