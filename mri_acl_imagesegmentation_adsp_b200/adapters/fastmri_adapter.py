"""The data format on the INPUT side of the path: fastMRI ``.h5`` volumes.

Twin of ``REF/src/adapters/fastmri_adapter.py:4-52`` (``FastMRISinglecoilAdapter``): ``discover_records`` lists one
``{'filepath', 'slice_idx'}`` per slice of every ``*.h5`` under the root, ``load_record`` returns the record dict the
preprocessor consumes (``'kspace'`` = ``hf['kspace'][s]``, ``'target'`` = the first of ``reconstruction_rss`` /
``reconstruction_esc`` / ``reconstruction`` present, ``'meta'``).  The reference never reads the file's own ``mask`` dataset;
``load_record(..., with_sampling_mask=True)`` adds it as ``'sampling_mask'`` (the name avoids the record's ``'mask'`` key, which
is a segmentation mask) for the multicoil stage.

HDF5 access goes through ``opener(path)``: a context manager whose value behaves like ``h5py.File`` for the three things the
adapter does (``name in f``, ``f[name].shape``, ``f[name][s]``).  The default opener is ``h5py.File`` and fails loudly
where h5py is not installed; ``NpzVolumeFile`` is the HDF5-free stand-in the tests use (same mapping interface over a numpy
archive stored under the ``.h5`` name).
"""
from __future__ import annotations

import glob
import os
from typing import Any, Callable, Dict, List, Optional

import numpy as np


class NpzVolumeFile:
    """``h5py.File``-shaped view of an ``np.savez`` archive: ``with NpzVolumeFile(path) as f: f['kspace'][3]``."""

    def __init__(self, path: str, mode: str = "r"):
        if mode != "r":
            raise ValueError("read-only")
        self._z = np.load(path, allow_pickle=False)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self._z.close()
        return False

    def __contains__(self, name: str) -> bool:
        return name in self._z.files

    def __getitem__(self, name: str) -> np.ndarray:
        return self._z[name]

    def keys(self):
        return list(self._z.files)

    @staticmethod
    def write(path: str, **datasets: np.ndarray) -> None:
        with open(path, "wb") as f:            # (np.savez would append ".npz" to a name that lacks it)
            np.savez(f, **datasets)


def _h5py_opener(path: str):
    try:
        import h5py
    except ImportError as e:                   # loud: there is no silent fallback format
        raise ImportError("reading fastMRI .h5 files needs h5py; pass opener=NpzVolumeFile for numpy archives") from e
    return h5py.File(path, "r")


class FastMRISinglecoilAdapter:
    TARGET_KEYS = ("reconstruction_rss", "reconstruction_esc", "reconstruction")

    def __init__(self, root_dir: Optional[str] = None, env_key: str = "FASTMRI_ROOT", opener: Optional[Callable[[str], Any]] = None):
        resolved = root_dir or os.getenv(env_key)
        if not resolved:
            raise ValueError(f"Must provide root_dir or set env {env_key}")
        self.root_dir = resolved
        self.opener = opener or _h5py_opener

    def discover_records(self, root_dir: Optional[str] = None) -> List[Dict[str, Any]]:
        root = root_dir or self.root_dir
        if not root:
            raise ValueError("Missing root directory for fastMRI adapter")
        records = []
        for fp in sorted(glob.glob(os.path.join(root, "*.h5"))):
            with self.opener(fp) as hf:
                num_slices = hf["kspace"].shape[0]
            records.extend({"filepath": fp, "slice_idx": s} for s in range(num_slices))
        return records

    def load_record(self, record: Dict[str, Any], with_sampling_mask: bool = False) -> Dict[str, Any]:
        fp, s = record["filepath"], record["slice_idx"]
        target, target_key, smask = None, None, None
        with self.opener(fp) as hf:
            kspace = np.asarray(hf["kspace"][s])
            for cand in self.TARGET_KEYS:
                if cand in hf:
                    target = np.asarray(hf[cand][s])
                    target_key = cand
                    break
            if with_sampling_mask and "mask" in hf:
                smask = np.asarray(hf["mask"][()] if hasattr(hf["mask"], "shape") and hf["mask"].shape == () else hf["mask"][...])
        out = {"image": None, "mask": None, "label": None, "kspace": kspace, "target": target,
               "meta": {"filepath": fp, "slice_idx": s, "dataset": "fastmri", "target_key": target_key,
                        "adapter": "fastmri_singlecoil-h5"}}
        if with_sampling_mask:
            out["sampling_mask"] = smask
        return out

    def load_volume_kspace(self, filepath: str) -> np.ndarray:
        """All slices of one file at once, ``(S, [C,] H, W)`` complex64: one contiguous host array for the fused stage
        (``recon.pipeline.HostPipeline`` / ``zero_filled_rss``) instead of one record per slice."""
        with self.opener(filepath) as hf:
            return np.ascontiguousarray(hf["kspace"][...], dtype=np.complex64)
