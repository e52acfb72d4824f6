"""Twin of the vendored GRAPPA class (``ZIP!/fastmri_prostate/reconstruction/grappa.py``), with the weight APPLICATION on
the GPU (the vendored README's stated bottleneck).

Same constructor, attributes and methods as the reference:

* ``Grappa(kspace, kernel_size=(5, 5), coil_axis=-1)`` extracts the unique kernel geometries of one k-space slice
  (``get_kernel_geometries``, ``:15-102``) -- host index work, done once per average and shared by all its slices;
* ``compute_weights(calib)`` (``:104-171``) solves the regularised normal equations per geometry with numpy on the host
  (a few small Hermitian systems per slice; "cuSOLVER-free CPU" in the survey's words);
* ``apply_weights(kspace, weights)`` (``:173-222``) fills every hole with ``W[ii] @ S`` -- ONE C-ABI call
  (``mriacl_grappa_apply_c64``); ``apply_weights_batch`` does the same for a stack of slices with their own weights in one
  launch, on the file's own axis order (no transposes).

Geometry keys (``patch_indices``) are the row numbers of ``np.unique(patches, axis=0)`` exactly as in the reference, so a
weights dict computed by either implementation can be applied by the other.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _device as D

_ITEM_HOLES = 256          # holes per work item of the kernel (grappa_kernels.cuh: GR_T * GR_HPT)


def _windows2d(a: np.ndarray, kx: int, ky: int) -> np.ndarray:
    """``skimage.util.view_as_windows(a, (kx, ky))`` for a 2-D array: all overlapping windows, step 1."""
    return np.lib.stride_tricks.sliding_window_view(a, (kx, ky))


class Grappa:
    def __init__(self, kspace: Any, kernel_size: Tuple[int, int] = (5, 5), coil_axis: int = -1) -> None:
        if isinstance(kspace, torch.Tensor):
            kspace = kspace.detach().cpu().numpy()
        self.kspace = np.asarray(kspace)
        if self.kspace.ndim != 3:
            raise ValueError(f"kspace must be 3-D (two k-space axes and a coil axis), got {self.kspace.shape}")
        self.kernel_size = tuple(kernel_size)
        self.coil_axis = coil_axis
        self.lamda = 0.01
        self._plan = None
        self.kernel_var_dict = self.get_kernel_geometries()

    # ------------------------------------------------------------------------------------------- geometries (host)
    def get_kernel_geometries(self):
        """``grappa.py:15-102``.  Returns the same dict (``patches`` tiled over coils, ``patch_indices``, ``holes_x``,
        ``holes_y`` in padded coordinates); like the reference it returns the k-space itself when there are no holes."""
        self.kspace = np.moveaxis(self.kspace, self.coil_axis, -1)
        if np.sum((np.abs(self.kspace[..., 0]) == 0).flatten()) == 0:
            return np.moveaxis(self.kspace, -1, self.coil_axis)
        kx, ky = self.kernel_size
        kx2, ky2 = int(kx / 2), int(ky / 2)
        nc = self.kspace.shape[-1]
        self.kspace = np.pad(self.kspace, ((kx2, kx2), (ky2, ky2), (0, 0)), mode="constant")
        mask = np.ascontiguousarray(np.abs(self.kspace[..., 0]) > 0)
        win = _windows2d(mask, kx, ky)                          # (X, Y, kx, ky)
        psh = win.shape[:2]
        # np.unique(P, axis=0) sorts the flattened boolean patches lexicographically (False < True): packing a patch into
        # an integer with its FIRST element as the most significant bit gives the same order and the same row numbers
        codes = np.zeros(psh, dtype=np.int64)
        for i in range(kx):
            for j in range(ky):
                codes = (codes << 1) | win[:, :, i, j]
        ucodes, iidx = np.unique(codes.reshape(-1), return_inverse=True)
        P = ((ucodes[:, None] >> np.arange(kx * ky - 1, -1, -1)[None, :]) & 1).astype(bool).reshape(-1, kx, ky)
        validP = np.argwhere(~P[:, kx2, ky2]).squeeze()
        invalidP = np.argwhere(np.all(P == 0, axis=(1, 2)))
        validP = np.setdiff1d(validP, invalidP, assume_unique=True)
        validP = np.atleast_1d(validP)
        holes_x, holes_y = {}, {}
        iidx = iidx.reshape(-1)
        order = np.argsort(iidx, kind="stable")
        bounds = np.searchsorted(iidx[order], np.arange(len(ucodes) + 1))
        for ii in validP:
            flat = order[bounds[ii]:bounds[ii + 1]]             # ascending flat index = the reference's argwhere order
            x, y = np.unravel_index(flat, psh)
            holes_x[ii] = np.atleast_1d(x + kx2)
            holes_y[ii] = np.atleast_1d(y + ky2)
        self._P2d = P
        return {"patches": np.tile(P[..., None], (1, 1, 1, nc)), "patch_indices": validP, "holes_x": holes_x, "holes_y": holes_y}

    # ------------------------------------------------------------------------------------------- weights (host)
    def compute_weights(self, calib: np.ndarray) -> Dict[int, np.ndarray]:
        """``grappa.py:104-171``: per geometry ``W = ((S^H S + lamda0 I)^-1 S^H T)^T`` over all windows of the calibration
        data, ``lamda0 = 0.01 * ||S^H S|| / n``.  numpy on the host, in the calibration data's own precision."""
        if isinstance(calib, torch.Tensor):
            calib = calib.detach().cpu().numpy()
        calib = np.moveaxis(np.asarray(calib), self.coil_axis, -1)
        kx, ky = self.kernel_size
        kx2, ky2 = int(kx / 2), int(ky / 2)
        nc = calib.shape[-1]
        calib = np.pad(calib, ((kx2, kx2), (ky2, ky2), (0, 0)), mode="constant")
        A = np.lib.stride_tricks.sliding_window_view(calib, (kx, ky, nc)).reshape((-1, kx, ky, nc))
        weights = {}
        for ii in self.kernel_var_dict["patch_indices"]:
            S = A[:, self.kernel_var_dict["patches"][ii, ...]]
            T = A[:, kx2, ky2, :]
            ShS = S.conj().T @ S
            ShT = S.conj().T @ T
            lamda0 = self.lamda * np.linalg.norm(ShS) / ShS.shape[0]
            weights[ii] = np.linalg.solve(ShS + lamda0 * np.eye(ShS.shape[0]), ShT).T
        return weights

    # ------------------------------------------------------------------------------------------- apply (device)
    def _device_plan(self, dev: torch.device, lanes_along_x: bool = False):
        """Kernel tables of this object's geometries: holes grouped by geometry (unpadded coordinates), work items of at most
        256 holes, window offsets of the sampled positions, and where each geometry's weights sit in a slice's block.
        ``lanes_along_x``: order the holes of a geometry y-major so that consecutive holes (the lanes of a warp) are neighbours
        along the FIRST kernel axis -- chosen when that axis has the smaller memory stride, so a warp's source loads fall
        into a few cache lines instead of 32."""
        if self._plan is not None and self._plan["dev"] == dev and self._plan["lanes_along_x"] == lanes_along_x:
            return self._plan
        kv = self.kernel_var_dict
        kx, ky = self.kernel_size
        kx2, ky2 = int(kx / 2), int(ky / 2)
        X, Y = self.kspace.shape[0] - 2 * kx2, self.kspace.shape[1] - 2 * ky2
        nc = self.kspace.shape[-1]
        geoms = [int(g) for g in kv["patch_indices"]]
        hole_xy, item_geom, item_first, item_count = [], [], [], []
        src_start, src_off, w_start = [0], [], []
        w_total, n_holes, max_src = 0, 0, 0
        for gi, g in enumerate(geoms):
            pat = self._P2d[g]
            ii, jj = np.nonzero(pat)                              # row-major: the order boolean indexing flattens S in
            src_off.extend((ii * 8 + jj).tolist())
            src_start.append(len(src_off))
            max_src = max(max_src, len(ii))
            w_start.append(w_total)
            w_total += nc * len(ii) * nc
            hx, hy = (kv["holes_x"][g] - kx2).astype(np.int64), (kv["holes_y"][g] - ky2).astype(np.int64)
            if lanes_along_x:
                order = np.lexsort((hx, hy))                      # y-major: consecutive entries step along x
                hx, hy = hx[order], hy[order]
            xy = hx * Y + hy
            hole_xy.append(xy)
            for f in range(0, len(xy), _ITEM_HOLES):
                item_geom.append(gi)
                item_first.append(n_holes + f)
                item_count.append(min(_ITEM_HOLES, len(xy) - f))
            n_holes += len(xy)
        i32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int32), device=dev)
        self._plan = {"dev": dev, "lanes_along_x": lanes_along_x, "X": X, "Y": Y, "nc": nc, "geoms": geoms, "n_items": len(item_geom), "max_src": max_src,
                      "w_total": w_total, "w_start_host": w_start,
                      "hole_xy": i32(np.concatenate(hole_xy) if hole_xy else np.zeros(0)), "item_geom": i32(item_geom),
                      "item_first": i32(item_first), "item_count": i32(item_count), "src_start": i32(src_start),
                      "src_off": i32(src_off), "w_start": torch.as_tensor(np.asarray(w_start, dtype=np.int64), device=dev)}
        return self._plan

    def _pack_weights(self, weights_list: Sequence[Dict[int, np.ndarray]], plan) -> torch.Tensor:
        block = np.zeros((len(weights_list), max(1, plan["w_total"])), dtype=np.complex64)
        for s, wd in enumerate(weights_list):
            for gi, g in enumerate(plan["geoms"]):
                w = np.asarray(wd[g])
                block[s, plan["w_start_host"][gi]:plan["w_start_host"][gi] + w.size] = w.reshape(-1)
        return torch.from_numpy(block).to(plan["dev"])

    def apply_weights_batch(self, kspace: Any, weights_list: Sequence[Dict[int, np.ndarray]], axes: Tuple[int, int, int]) -> Any:
        """``kspace``: complex ``(S, d0, d1, d2)``; ``axes = (x_axis, y_axis, coil_axis)`` says which of the three trailing
        axes (numbered 0..2) is the first kernel axis, the second kernel axis and the coil axis.  Slice ``s`` is filled with
        ``weights_list[s]``.  Returns the filled k-space in the input's layout and type (complex64 arithmetic)."""
        if isinstance(self.kernel_var_dict, np.ndarray):        # no holes: the reference's apply would fail; nothing to fill
            return kspace
        mv = D.to_device_complex(kspace)
        k = mv.tensor.clone() if mv.tensor.data_ptr() == getattr(kspace, "data_ptr", lambda: 0)() else mv.tensor
        if k.ndim != 4 or len(weights_list) != k.shape[0] or sorted(axes) != [0, 1, 2]:
            raise ValueError("kspace must be (S, d0, d1, d2) with one weights dict per slice and axes a permutation of (0, 1, 2)")
        dims = k.shape[1:]
        strides = (dims[1] * dims[2], dims[2], 1)
        plan = self._device_plan(k.device, lanes_along_x=strides[axes[0]] < strides[axes[1]])
        if (dims[axes[0]], dims[axes[1]], dims[axes[2]]) != (plan["X"], plan["Y"], plan["nc"]):
            raise ValueError(f"k-space {tuple(dims)} with axes {axes} does not match the geometry plan "
                             f"({plan['X']}, {plan['Y']}, {plan['nc']})")
        if plan["n_items"]:
            W = self._pack_weights(weights_list, plan)
            D.lib().grappa_apply(k.data_ptr(), dims[0] * dims[1] * dims[2], strides[axes[0]], strides[axes[1]], strides[axes[2]],
                                 k.shape[0], plan["X"], plan["Y"], plan["nc"], self.kernel_size[0], self.kernel_size[1],
                                 plan["hole_xy"].data_ptr(), plan["n_items"], plan["item_geom"].data_ptr(),
                                 plan["item_first"].data_ptr(), plan["item_count"].data_ptr(), plan["src_start"].data_ptr(),
                                 plan["src_off"].data_ptr(), plan["max_src"], plan["w_start"].data_ptr(), W.data_ptr(),
                                 W.shape[1], D.stream_ptr())
            W.record_stream(torch.cuda.current_stream())
        return mv.back(k, widen=True)

    def apply_weights(self, kspace: Any, weights: Dict[int, np.ndarray]) -> Any:
        """``grappa.py:173-222``: one slice, coil axis = ``self.coil_axis``; result has the input's shape."""
        nd = 3
        ca = self.coil_axis % nd
        rest = [a for a in range(nd) if a != ca]
        is_np = isinstance(kspace, np.ndarray)
        out = self.apply_weights_batch(kspace[None] if is_np else kspace.unsqueeze(0), [weights], (rest[0], rest[1], ca))
        return out[0]
