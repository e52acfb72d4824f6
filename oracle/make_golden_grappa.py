"""Freeze outputs of the UNMODIFIED vendored GRAPPA class into ``tests/golden/grappa_vectors.npz``.

Run in the build container (needs ``/root/reference``)::

    python -m oracle.make_golden_grappa

``ZIP!/fastmri_prostate/reconstruction/grappa.py`` is imported as it is (``oracle/ref_shim.py``; its one third-party import,
``skimage.util.view_as_windows``, is supplied by numpy's identical ``sliding_window_view``).  Per case of
``synth.GRAPPA_CASES``: the geometry keys, the number of holes per geometry, the weights of every geometry and the filled
k-space (``Grappa(...).compute_weights(calib)`` -> ``apply_weights``), plus the SENSE-style combine of the DWI chain.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mri_acl_imagesegmentation_adsp_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "grappa_vectors.npz")


def main() -> None:
    if not ref_shim.available():
        raise SystemExit("reference checkout not found; golden vectors can only be made in the build container")
    G = ref_shim.prostate().grappa.Grappa
    vec, man = {}, {"numpy": np.__version__, "cases": {}}
    for name, pe, nc, ro, acc, acs, cal, seed in synth.GRAPPA_CASES:
        k, calib = synth.grappa_case_inputs(name)
        g = G(k.copy(), kernel_size=(5, 5), coil_axis=1)
        w = g.compute_weights(calib.copy())
        filled = g.apply_weights(k.copy(), w)
        keys = [int(i) for i in g.kernel_var_dict["patch_indices"]]
        vec[f"{name}/patch_indices"] = np.asarray(keys, dtype=np.int64)
        vec[f"{name}/holes_per_geometry"] = np.asarray([len(g.kernel_var_dict["holes_x"][i]) for i in keys], dtype=np.int64)
        for i in keys:
            vec[f"{name}/weights_{i}"] = np.asarray(w[i]).astype(np.complex64)      # (the reference solves in complex128)
        sub = 1 if filled.size < 50000 else 2
        vec[f"{name}/filled"] = np.ascontiguousarray(filled[:, :, ::sub])
        man["cases"][name] = {"shape": [pe, nc, ro], "coil_axis": 1, "acceleration": acc, "acs": acs, "calib_lines": cal,
                              "seed": seed, "ro_subsample": sub, "filled_dtype": str(filled.dtype)}
    # SENSE-style combine (dwi/prostate_dwi_recon.py:106-109), spelled as the reference spells it
    img = synth.gaussian_kspace((3, 5, 12, 10), 611)
    sens = synth.gaussian_kspace((3, 5, 12, 10), 612)
    vec["sense/abs_sum"] = np.abs(np.sum(img * (sens.conj()), axis=1)).astype(np.float32)
    # the whole T2 reconstruction (prostate_t2_recon.py:9-78) on a small three-average volume, header parsing included
    k, calib, hdr = synth.t2_recon_case_inputs()
    rec = ref_shim.prostate().t2.t2_reconstruction(k.copy(), calib.copy(), hdr)["reconstruction_rss"]
    vec["t2_recon/reconstruction_rss_sub2"] = rec[:, ::2, ::2].astype(np.float32)
    vec["t2_recon/padding"] = np.array([ref_shim.prostate().mri_data.get_padding(hdr)], dtype=np.float64)
    man["t2_recon"] = {"shape": list(synth.T2_RECON_CASE[0]), "out": list(rec.shape), "dtype": str(rec.dtype)}
    np.savez_compressed(OUT, **vec)
    with open(os.path.join(ROOT, "tests", "golden", "grappa_manifest.json"), "w") as f:
        json.dump(man, f, indent=1)
    print(f"wrote {OUT}: {len(vec)} arrays, {os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
