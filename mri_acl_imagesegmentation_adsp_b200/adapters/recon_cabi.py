"""ctypes binding of ``libmriacl_recon.so`` (C ABI in ``include/mriacl_recon.h``).

This is the thin C-ABI layer the north star places in ``src/adapters``: raw pointers, sizes and
a CUDA stream handle go in; status codes come out and are turned into the reference's error
convention (``ValueError`` for shape/argument violations; ``RuntimeError``
for CUDA failures).  There is no CPU fallback: if the shared library has not been built, or a
symbol the header declares is missing, loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

ABI_VERSION = 1

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3, -4
FLIP_ROWS, NORM_INSTANCE, FORCE_GENERIC, SEQUENTIAL = 0x1, 0x2, 0x4, 0x8
SCHED_FUSED, SCHED_OVERLAP = 0x10, 0x20    # experimental kernel schedules of the fused plan
SCHED_CORESIDENT = 0x80                    # one persistent kernel: column team + row team on every SM, normalise fused
SCHED_PIPELINED = 0x800                    # small chunks on two streams, T kept in L2
SCHED_PAIR = 0x40                          # column pass -> pair row pass (rowpair.cuh) -> normalise
ONLY_COLPASS, ONLY_ROWPASS, ONLY_NORM = 0x100, 0x200, 0x400   # profiling: single phases of the fused plan
PACKED_COLUMNS = 0x1000                    # k-space holds the sampled columns only (mriacl_pack_columns_host)
PATH_NONE, PATH_GENERIC, PATH_FUSED = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
#: the product library; MRIACL_RECON_LIBRARY selects another build of the same C ABI (A/B runs of the experimental
#: schedules: `make -C csrc experimental` -> libmriacl_recon_exp.so).  Never a CPU library: there is none.
DEFAULT_LIBRARY = os.environ.get("MRIACL_RECON_LIBRARY") or os.path.join(os.path.dirname(_HERE), "csrc", "libmriacl_recon.so")

_vp, _i, _u, _f, _ll, _sz = C.c_void_p, C.c_int, C.c_uint, C.c_float, C.c_longlong, C.c_size_t
_fp = C.POINTER(C.c_float)

# name -> (restype, argtypes); mirrors include/mriacl_recon.h one to one
SIGNATURES = {
    "mriacl_abi_version": (_i, []),
    "mriacl_last_error": (C.c_char_p, []),
    "mriacl_supported": (_i, [_i, _i]),
    "mriacl_launch_count": (C.c_uint64, []),
    "mriacl_recon_rss_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i, _i, _i, _fp, _u]),
    "mriacl_recon_rss_f32": (_i, [_vp, _ll, _ll, _fp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _u, _f,
                                   _vp, _sz, _vp]),
    "mriacl_pack_columns_host": (_i, [_vp, _vp, _ll, _i, _fp, _i]),
    "mriacl_ifft2c_abs_workspace_bytes": (_sz, [_i, _i, _i]),
    "mriacl_ifft2c_abs_f32": (_i, [_vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "mriacl_fft2c_c64": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "mriacl_complex_abs_f32": (_i, [_vp, _vp, _sz, _i, _vp]),
    "mriacl_rss_f32": (_i, [_vp, _vp, _sz, _i, _sz, _i, _vp]),
    "mriacl_center_crop_or_pad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "mriacl_normalize_instance_f32": (_i, [_vp, _vp, _vp, _i, _sz, _f, _vp]),
    "mriacl_percentile_clip_f32": (_i, [_vp, _vp, _vp, _i, _sz, _f, _f, _vp]),
    "mriacl_resize_bilinear_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "mriacl_resize_mask_u8": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "mriacl_zscore_preview_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _sz, _vp]),
    "mriacl_grappa_apply_c64": (_i, [_vp, _ll, _ll, _ll, _ll, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i,
                                     _vp, _vp, _ll, _vp]),
    "mriacl_stack25d_f32": (_i, [_vp, _vp, _i, _sz, _i, _i, _vp, _vp, _vp]),
    "mriacl_sense_combine": (_i, [_vp, _vp, _vp, _i, _i, _sz, _i, _i, _vp]),
    "mriacl_clip_resize_zscore_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _vp]),
}


class ReconLibraryError(RuntimeError):
    pass


class ReconLibrary:
    """One loaded ``libmriacl_recon.so``.  All pointer arguments are plain integers
    (``tensor.data_ptr()``); the mask is a host ``numpy`` float32 vector or ``None``."""

    def __init__(self, path: str = DEFAULT_LIBRARY):
        if not os.path.isfile(path):
            raise ReconLibraryError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C mri_acl_imagesegmentation_adsp_b200/csrc`). There is no CPU fallback.")
        self.path = path
        self._lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(self._lib, name)
            except AttributeError as e:
                raise ReconLibraryError(f"{path} does not export {name}") from e
            fn.restype, fn.argtypes = res, args
        got = self._lib.mriacl_abi_version()
        if got != ABI_VERSION:
            raise ReconLibraryError(f"ABI version {got} != expected {ABI_VERSION}")

    # -- helpers -----------------------------------------------------------------------
    def last_error(self) -> str:
        msg = self._lib.mriacl_last_error()
        return msg.decode("utf-8", "replace") if msg else ""

    def _check(self, rc: int) -> None:
        if rc == OK:
            return
        msg = self.last_error()
        if rc in (ERR_INVALID, ERR_UNSUPPORTED):
            raise ValueError(msg)
        raise RuntimeError(f"libmriacl_recon error {rc}: {msg}")

    @staticmethod
    def _mask_ptr(mask):
        if mask is None:
            return None, None
        import numpy as np
        m = np.ascontiguousarray(mask, dtype=np.float32)
        return m, m.ctypes.data_as(_fp)

    # -- queries -----------------------------------------------------------------------
    def supported(self, h: int, w_padded: int) -> int:
        return int(self._lib.mriacl_supported(h, w_padded))

    def launch_count(self) -> int:
        return int(self._lib.mriacl_launch_count())

    def recon_rss_workspace_bytes(self, slices, a, c, h, w, pad_left, w_padded, out_h, out_w, mask=None, flags=0) -> int:
        keep, mp = self._mask_ptr(mask)
        n = int(self._lib.mriacl_recon_rss_workspace_bytes(slices, a, c, h, w, pad_left, w_padded, out_h, out_w, mp, flags))
        if n == 0:
            raise ValueError(self.last_error())
        return n

    def ifft2c_abs_workspace_bytes(self, b, h, w) -> int:
        return int(self._lib.mriacl_ifft2c_abs_workspace_bytes(b, h, w))

    # -- compute -----------------------------------------------------------------------
    def recon_rss(self, kspace_ptr, slice_stride, avg_stride, mask, out_ptr, mean_std_ptr, b, a, c, h, w,
                  pad_left, w_padded, out_h, out_w, flags, eps, workspace_ptr, workspace_bytes, stream=0) -> None:
        keep, mp = self._mask_ptr(mask)
        self._check(self._lib.mriacl_recon_rss_f32(kspace_ptr, slice_stride, avg_stride, mp, out_ptr, mean_std_ptr or None,
                                                    b, a, c, h, w, pad_left, w_padded, out_h, out_w, flags, eps,
                                                    workspace_ptr, workspace_bytes, stream or None))

    def pack_columns_host(self, src_host_ptr, dst_host_ptr, n_rows, w, mask, n_threads=0) -> int:
        """Host gather of the sampled columns (both pointers are HOST addresses); returns n_act.  ctypes drops the
        GIL for the duration of the call."""
        keep, mp = self._mask_ptr(mask)
        n = int(self._lib.mriacl_pack_columns_host(src_host_ptr, dst_host_ptr, n_rows, w, mp, n_threads))
        if n < 0:
            self._check(n)
        return n

    def ifft2c_abs(self, k_ptr, out_ptr, b, h, w, workspace_ptr, workspace_bytes, stream=0) -> None:
        self._check(self._lib.mriacl_ifft2c_abs_f32(k_ptr, out_ptr, b, h, w, workspace_ptr, workspace_bytes, stream or None))

    def fft2c(self, in_ptr, out_ptr, b, h, w, inverse: bool, stream=0) -> None:
        self._check(self._lib.mriacl_fft2c_c64(in_ptr, out_ptr, b, h, w, 1 if inverse else 0, stream or None))

    def complex_abs(self, in_ptr, out_ptr, n, squared: bool, stream=0) -> None:
        self._check(self._lib.mriacl_complex_abs_f32(in_ptr, out_ptr, n, 1 if squared else 0, stream or None))

    def rss(self, in_ptr, out_ptr, outer, c, inner, is_complex: bool, stream=0) -> None:
        self._check(self._lib.mriacl_rss_f32(in_ptr, out_ptr, outer, c, inner, 1 if is_complex else 0, stream or None))

    def center_crop_or_pad(self, in_ptr, out_ptr, b, h, w, out_h, out_w, elem_bytes, stream=0) -> None:
        self._check(self._lib.mriacl_center_crop_or_pad(in_ptr, out_ptr, b, h, w, out_h, out_w, elem_bytes, stream or None))

    def normalize_instance(self, in_ptr, out_ptr, mean_std_ptr, b, n, eps, stream=0) -> None:
        self._check(self._lib.mriacl_normalize_instance_f32(in_ptr, out_ptr, mean_std_ptr or None, b, n, eps, stream or None))


    # -- steps after the reconstruction (mri_preprocess.py:182-191,216-233) ---------------
    def percentile_clip(self, in_ptr, out_ptr, lo_hi_ptr, b, n, pmin, pmax, stream=0) -> None:
        self._check(self._lib.mriacl_percentile_clip_f32(in_ptr, out_ptr or None, lo_hi_ptr or None, b, n, pmin, pmax, stream or None))

    def resize_bilinear(self, in_ptr, out_ptr, b, h, w, oh, ow, stream=0) -> None:
        self._check(self._lib.mriacl_resize_bilinear_f32(in_ptr, out_ptr, b, h, w, oh, ow, stream or None))

    def resize_mask(self, in_ptr, out_ptr, b, h, w, oh, ow, stream=0) -> None:
        self._check(self._lib.mriacl_resize_mask_u8(in_ptr, out_ptr, b, h, w, oh, ow, stream or None))

    def zscore_preview(self, in_ptr, mask_ptr, z_ptr, p01_ptr, stats_ptr, b, n, stream=0) -> None:
        self._check(self._lib.mriacl_zscore_preview_f32(in_ptr, mask_ptr or None, z_ptr or None, p01_ptr or None,
                                                        stats_ptr or None, b, n, stream or None))

    def clip_resize_zscore(self, img_ptr, mask_ptr, z_ptr, p01_ptr, out_mask_ptr, lo_hi_ptr, stats_ptr, b, h, w, oh, ow,
                           pmin, pmax, stream=0) -> None:
        self._check(self._lib.mriacl_clip_resize_zscore_f32(img_ptr, mask_ptr or None, z_ptr, p01_ptr or None,
                                                            out_mask_ptr or None, lo_hi_ptr, stats_ptr or None,
                                                            b, h, w, oh, ow, pmin, pmax, stream or None))


    def grappa_apply(self, k_ptr, slice_stride, sx, sy, sc, n_slices, x, y, c, kx, ky, hole_xy_ptr, n_items, item_geom_ptr,
                     item_first_ptr, item_count_ptr, geom_src_start_ptr, src_off_ptr, max_sources, geom_w_start_ptr,
                     weights_ptr, weights_per_slice, stream=0) -> None:
        self._check(self._lib.mriacl_grappa_apply_c64(k_ptr, slice_stride, sx, sy, sc, n_slices, x, y, c, kx, ky, hole_xy_ptr,
                                                      n_items, item_geom_ptr, item_first_ptr, item_count_ptr,
                                                      geom_src_start_ptr, src_off_ptr, max_sources, geom_w_start_ptr,
                                                      weights_ptr, weights_per_slice, stream or None))

    def stack25d(self, in_ptr, out_ptr, s, n, k, repeat: bool, mean_ptr, std_ptr, stream=0) -> None:
        self._check(self._lib.mriacl_stack25d_f32(in_ptr, out_ptr, s, n, k, 1 if repeat else 0, mean_ptr or None, std_ptr or None,
                                                  stream or None))

    def sense_combine(self, img_ptr, sens_ptr, out_ptr, b, c, n, shared_sens: bool, magnitude: bool, stream=0) -> None:
        self._check(self._lib.mriacl_sense_combine(img_ptr, sens_ptr, out_ptr, b, c, n, 1 if shared_sens else 0,
                                                   1 if magnitude else 0, stream or None))


_lock = threading.Lock()
_default: Optional[ReconLibrary] = None


def library() -> ReconLibrary:
    """The process-wide product library (``csrc/libmriacl_recon.so``); raises if it is not built."""
    global _default
    with _lock:
        if _default is None:
            _default = ReconLibrary(DEFAULT_LIBRARY)
        return _default
