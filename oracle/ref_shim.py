"""Import the UNMODIFIED reference functions for this path (build container only).

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` does not exist on the GPU box, so
nothing that runs there may import this module; it is used by
``oracle/make_golden.py`` (to freeze reference outputs into ``tests/golden/``) and
by ``tests/test_oracle_vs_live_reference.py`` (skipped when the checkout is absent).

Recipe: ``src/utils/kspace.py`` imports as-is; the
preprocess module needs ``skimage`` stubbed; the vendored ZIP's files are loaded by
path under a synthetic package because its package ``__init__`` pulls in ``h5py``.
No reference source is copied: the ZIP is extracted to a temp dir at run time.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import tempfile
import types
import zipfile

REF_ROOT = os.environ.get("MRIACL_REFERENCE_ROOT", "/root/reference")
_ZIP = os.path.join(REF_ROOT, "reference", "fastMRI_prostate-main.zip")
_cache: dict = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "utils", "kspace.py")) and os.path.isfile(_ZIP)


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__dict__["__mriacl_stub__"] = True
        sys.modules[name] = mod
    for k, v in attrs.items():
        if not hasattr(mod, k):
            setattr(mod, k, v)
    return mod


def _zip_root() -> str:
    if "zip" not in _cache:
        d = tempfile.mkdtemp(prefix="mriacl_refzip_")
        with zipfile.ZipFile(_ZIP) as z:
            z.extractall(d)
        _cache["zip"] = os.path.join(d, "fastMRI_prostate-main")
    return _cache["zip"]


def _load_by_path(qualname: str, path: str, package: str | None = None):
    spec = importlib.util.spec_from_file_location(qualname, path)
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[qualname] = mod
    spec.loader.exec_module(mod)
    return mod


def kspace_utils():
    """REF/src/utils/kspace.py -> module with fft2c, ifft2c, complex_abs, center_crop_or_pad."""
    if "kspace" not in _cache:
        _cache["kspace"] = _load_by_path("_mriacl_ref_kspace", os.path.join(REF_ROOT, "src", "utils", "kspace.py"))
    return _cache["kspace"]


def knee_preprocessor_cls():
    """REF/src/preprocess/mri_preprocess.py::MRIKneePreprocessor (skimage stubbed)."""
    if "pre" not in _cache:
        _stub("skimage")
        _stub("skimage.filters", threshold_otsu=None)
        _stub("skimage.morphology", remove_small_objects=None, binary_closing=None, binary_opening=None,
              disk=None, remove_small_holes=None)
        _stub("skimage.restoration", denoise_nl_means=None, estimate_sigma=None)
        _stub("skimage.transform", resize=None)
        path = os.path.join(REF_ROOT, "src", "preprocess", "mri_preprocess.py")
        try:
            mod = _load_by_path("_mriacl_ref_preprocess", path)
        except ImportError:
            # the module imports a handful of names; make any missing one resolvable
            import re
            src = open(path, "r", encoding="utf-8").read()
            for m in re.finditer(r"^from\s+(skimage[\w.]*)\s+import\s+(.+)$", src, re.M):
                names = [n.strip().split(" as ")[0] for n in m.group(2).strip("() ").split(",") if n.strip()]
                _stub(m.group(1), **{n: None for n in names})
            mod = _load_by_path("_mriacl_ref_preprocess", path)
        _cache["pre"] = mod.MRIKneePreprocessor
    return _cache["pre"]


def fastmri_dl():
    """Vendored fastMRI torch functions: namespace with fftc, math_fn, coil_combine, transforms."""
    if "dl" not in _cache:
        root = os.path.join(_zip_root(), "DL_reconstruction")
        pkg = types.ModuleType("_mriacl_ref_dl")
        pkg.__path__ = [root]
        sys.modules["_mriacl_ref_dl"] = pkg
        ns = types.SimpleNamespace()
        for name in ("math_fn", "fftc", "coil_combine"):
            setattr(ns, name, _load_by_path(f"_mriacl_ref_dl.{name}", os.path.join(root, f"{name}.py"),
                                            package="_mriacl_ref_dl"))
        data_pkg = types.ModuleType("_mriacl_ref_dl.data")
        data_pkg.__path__ = [os.path.join(root, "data")]
        sys.modules["_mriacl_ref_dl.data"] = data_pkg
        ns.transforms = _load_by_path("_mriacl_ref_dl.data.transforms", os.path.join(root, "data", "transforms.py"),
                                      package="_mriacl_ref_dl.data")
        _cache["dl"] = ns
    return _cache["dl"]


def prostate():
    """Vendored numpy prostate functions: namespace with utils (ifftnd, center_crop_im),
    t2 (create_coil_combined_im, rss) and mri_data (get_padding, zero_pad_kspace_hdr)."""
    if "prostate" not in _cache:
        root = _zip_root()
        _stub("h5py")
        _stub("skimage")
        # skimage.util.view_as_windows(arr, shape) with the default step 1 IS numpy's sliding_window_view: the one
        # third-party function grappa.py needs is supplied by its numpy equivalent, nothing of the reference is changed
        import numpy as _np
        _stub("skimage.util", view_as_windows=None)
        sys.modules["skimage.util"].view_as_windows = lambda arr, window_shape, step=1: _np.lib.stride_tricks.sliding_window_view(arr, window_shape)
        if root not in sys.path:
            sys.path.insert(0, root)
        ns = types.SimpleNamespace()
        ns.utils = importlib.import_module("fastmri_prostate.reconstruction.utils")
        ns.mri_data = importlib.import_module("fastmri_prostate.data.mri_data")
        ns.t2 = importlib.import_module("fastmri_prostate.reconstruction.t2.prostate_t2_recon")
        ns.grappa = importlib.import_module("fastmri_prostate.reconstruction.grappa")
        _cache["prostate"] = ns
    return _cache["prostate"]


ISMRMRD_HEADER_TEMPLATE = """<?xml version="1.0"?>
<ismrmrdHeader xmlns="http://www.ismrm.org/ISMRMRD">
 <encoding>
  <encodedSpace><matrixSize><x>{enc_x}</x><y>{enc_y}</y><z>1</z></matrixSize></encodedSpace>
  <encodingLimits><kspace_encoding_step_1><minimum>0</minimum><maximum>{max_pe}</maximum><center>{center}</center></kspace_encoding_step_1></encodingLimits>
 </encoding>
</ismrmrdHeader>"""


def synthetic_header(enc_x: int, max_pe: int) -> str:
    """Six-line synthetic ISMRMRD header carrying only what ``get_padding`` reads."""
    return ISMRMRD_HEADER_TEMPLATE.format(enc_x=enc_x, enc_y=max_pe + 1, max_pe=max_pe, center=(max_pe + 1) // 2)
