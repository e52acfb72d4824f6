"""CPU (numpy) restatement of the reference's k-space -> image input stage.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  The product path never
imports this module; the CUDA library has no CPU fallback.

Every function cites the reference lines it restates.  ``REF`` = the read-only
checkout of bonhchi/mri_acl_imagesegmentation_adsp, ``ZIP!`` = a member of
``REF/reference/fastMRI_prostate-main.zip`` (the vendored cai2r/fastMRI_prostate
tree that holds the multicoil pieces of the path).

Pinning: the reference ships no tests, golden vectors or
fixtures for this path.  The pin is therefore made by this build:
``oracle/make_golden.py`` imports the reference's own functions (through
``oracle/ref_shim.py``) in the build container, runs them on seeded synthetic
k-space and freezes their outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those frozen
reference outputs.  Two pieces have no reference implementation at all and are
"parity unpinned" in the reference's own terms: the undersampling-mask
*generator* (`equispaced_mask`; no mask code exists in the reference) and the
FFT library result itself (numpy's pocketfft / torch.fft are third-party
dependencies, ``numpy>=2.0.0`` at REF/src/requirements.txt:20).  Mask
*application* and crop indexing are pinned exactly.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------
# a1/a2  centred 2-D FFT pair            REF/src/utils/kspace.py:4-16
# --------------------------------------------------------------------------

_AX = (-2, -1)


def fft2c(x: np.ndarray) -> np.ndarray:
    """Centred, orthonormal 2-D FFT over the last two axes (kspace.py:4-9)."""
    shifted = np.fft.ifftshift(x, axes=_AX)
    spec = np.fft.fft2(shifted, norm="ortho")
    return np.fft.fftshift(spec, axes=_AX)


def ifft2c(x: np.ndarray) -> np.ndarray:
    """Centred, orthonormal 2-D inverse FFT over the last two axes (kspace.py:11-16).

    complex64 in -> complex64 out on numpy >= 2 (pocketfft keeps single precision).
    """
    shifted = np.fft.ifftshift(x, axes=_AX)
    img = np.fft.ifft2(shifted, norm="ortho")
    return np.fft.fftshift(img, axes=_AX)


# --------------------------------------------------------------------------
# a4  magnitude                           REF/src/utils/kspace.py:18-20
#                                         ZIP!/DL_reconstruction/math_fn.py:55-86
# --------------------------------------------------------------------------


def complex_abs(x: np.ndarray) -> np.ndarray:
    """sqrt(re^2 + im^2) of a complex array (kspace.py:18-20)."""
    re, im = x.real, x.imag
    return np.sqrt(re * re + im * im)


def complex_abs_sq_ri(data: np.ndarray) -> np.ndarray:
    """re^2+im^2 on the fastMRI real-view layout (..., 2) (math_fn.py:72-86)."""
    if data.shape[-1] != 2:
        raise ValueError("Tensor does not have separate complex dim.")
    return (data ** 2).sum(axis=-1)


def complex_abs_ri(data: np.ndarray) -> np.ndarray:
    """|z| on the real-view layout (..., 2) (math_fn.py:55-69)."""
    return np.sqrt(complex_abs_sq_ri(data))


# --------------------------------------------------------------------------
# a3  single-coil magnitude recon         REF/src/preprocess/mri_preprocess.py:149-160,177-180
# --------------------------------------------------------------------------


def ifft2c_single(kspace_2d: np.ndarray) -> np.ndarray:
    """|ifft2c(k)| as float32 for ONE (H, W) complex slice.

    ValueError when ``ndim != 2`` (``_ensure_2d``, mri_preprocess.py:177-180).
    """
    if kspace_2d.ndim != 2:
        raise ValueError(f"kspace must be (H,W), got {kspace_2d.shape}")
    img = np.fft.fftshift(
        np.fft.ifft2(np.fft.ifftshift(kspace_2d, axes=_AX), norm="ortho"), axes=_AX
    )
    return np.abs(img).astype(np.float32)


# --------------------------------------------------------------------------
# a6  centre crop (three reference spellings, identical indices when n >= out)
# --------------------------------------------------------------------------


def crop_start(n: int, out: int) -> int:
    """First kept index of a centred window: (n - out) // 2.

    kspace.py:27-28, transforms.py:62-63 and utils.py:70-73 all reduce to this
    for ``n >= out`` (the float form ``int(n/2 - out/2)`` truncates x.5 down).
    """
    return (n - out) // 2


def center_crop_or_pad(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Crop or zero-pad the last two axes about the centre (kspace.py:22-31)."""
    h, w = img.shape[-2:]
    canvas = np.zeros(img.shape[:-2] + (out_h, out_w), dtype=img.dtype)
    keep_h, keep_w = min(h, out_h), min(w, out_w)
    sh, sw = (h - keep_h) // 2, (w - keep_w) // 2
    dh, dw = (out_h - keep_h) // 2, (out_w - keep_w) // 2
    canvas[..., dh:dh + keep_h, dw:dw + keep_w] = img[..., sh:sh + keep_h, sw:sw + keep_w]
    return canvas


def center_crop(data: np.ndarray, shape: Tuple[int, int]) -> np.ndarray:
    """fastMRI centre crop of the last two axes (transforms.py:45-67).

    ValueError("Invalid shapes.") when the crop exceeds the data (:59-60).
    """
    if not (0 < shape[0] <= data.shape[-2] and 0 < shape[1] <= data.shape[-1]):
        raise ValueError("Invalid shapes.")
    r0 = crop_start(data.shape[-2], shape[0])
    c0 = crop_start(data.shape[-1], shape[1])
    return data[..., r0:r0 + shape[0], c0:c0 + shape[1]]


def center_crop_im(im_3d: np.ndarray, crop_to_size: Sequence[int]) -> np.ndarray:
    """Prostate centre crop of (slices, y, x) (ZIP!/fastmri_prostate/reconstruction/utils.py:54-73)."""
    x0 = im_3d.shape[-1] / 2 - crop_to_size[0] / 2
    y0 = im_3d.shape[-2] / 2 - crop_to_size[1] / 2
    return im_3d[:, int(y0):int(crop_to_size[1] + y0), int(x0):int(crop_to_size[0] + x0)]


# --------------------------------------------------------------------------
# a5  root-sum-of-squares coil combine
# --------------------------------------------------------------------------


def rss(data: np.ndarray, dim: int = 0) -> np.ndarray:
    """sqrt(sum(data**2, dim)) for REAL data (ZIP!/DL_reconstruction/coil_combine.py:12-25)."""
    return np.sqrt((data ** 2).sum(axis=dim))


def rss_complex_ri(data: np.ndarray, dim: int = 0) -> np.ndarray:
    """RSS on the real-view layout (..., 2) (coil_combine.py:28-41)."""
    return np.sqrt(complex_abs_sq_ri(data).sum(axis=dim))


def rss_np(sig: np.ndarray, axis: int = -1) -> np.ndarray:
    """sqrt(sum(|sig|^2, axis)) for complex numpy data
    (ZIP!/fastmri_prostate/reconstruction/t2/prostate_t2_recon.py:105-121)."""
    return np.sqrt(np.sum(np.abs(sig) ** 2, axis))


# --------------------------------------------------------------------------
# a7  normalisation                        ZIP!/DL_reconstruction/data/transforms.py:120-162
# --------------------------------------------------------------------------


def normalize(data, mean, stddev, eps=0.0):
    """(data - mean) / (stddev + eps) (transforms.py:120-140)."""
    return (data - mean) / (stddev + eps)


def normalize_instance(data: np.ndarray, eps: float = 0.0):
    """Instance normalise with torch semantics: ``std`` is UNBIASED (N-1)
    (transforms.py:143-162: ``data.mean()``, ``data.std()``).

    Returns (normalised, mean, std); float32 in -> float32 out.
    """
    mean = data.mean(dtype=data.dtype)
    std = data.std(ddof=1, dtype=data.dtype)
    return normalize(data, mean, std, data.dtype.type(eps)), mean, std


# --------------------------------------------------------------------------
# a8  fastMRI real-view centred transforms  ZIP!/DL_reconstruction/fftc.py:14-65,93-165
# --------------------------------------------------------------------------


def to_real_view(data: np.ndarray) -> np.ndarray:
    """complex (...) -> real (..., 2) (``to_tensor``, transforms.py:14-29)."""
    if np.iscomplexobj(data):
        return np.stack((data.real, data.imag), axis=-1)
    return data


def _roll(x: np.ndarray, shifts: Sequence[int], dims: Sequence[int]) -> np.ndarray:
    """fftc.py:69-115 -- roll built from two narrows and a cat, one dim at a time."""
    if len(shifts) != len(dims):
        raise ValueError("len(shift) must match len(dim)")
    for s, d in zip(shifts, dims):
        n = x.shape[d]
        s %= n
        if s:
            head = np.take(x, range(0, n - s), axis=d)
            tail = np.take(x, range(n - s, n), axis=d)
            x = np.concatenate((tail, head), axis=d)
    return x


def _fftshift_ri(x, dims):   # fftc.py:118-140: shift n // 2
    return _roll(x, [x.shape[d] // 2 for d in dims], dims)


def _ifftshift_ri(x, dims):  # fftc.py:143-165: shift (n + 1) // 2
    return _roll(x, [(x.shape[d] + 1) // 2 for d in dims], dims)


def _c2ri(z: np.ndarray) -> np.ndarray:
    return np.stack((z.real, z.imag), axis=-1)


def ifft2c_new(data: np.ndarray) -> np.ndarray:
    """Centred ortho inverse FFT on (..., H, W, 2) float data (fftc.py:41-65)."""
    if data.shape[-1] != 2:
        raise ValueError("Tensor does not have separate complex dim.")
    data = _ifftshift_ri(data, [-3, -2])
    z = data[..., 0] + 1j * data[..., 1]
    z = np.fft.ifftn(z.astype(np.complex64, copy=False), axes=_AX, norm="ortho")
    return _fftshift_ri(_c2ri(z).astype(data.dtype, copy=False), [-3, -2])


def fft2c_new(data: np.ndarray) -> np.ndarray:
    """Forward twin of :func:`ifft2c_new` (fftc.py:14-38)."""
    if data.shape[-1] != 2:
        raise ValueError("Tensor does not have separate complex dim.")
    data = _ifftshift_ri(data, [-3, -2])
    z = data[..., 0] + 1j * data[..., 1]
    z = np.fft.fftn(z.astype(np.complex64, copy=False), axes=_AX, norm="ortho")
    return _fftshift_ri(_c2ri(z).astype(data.dtype, copy=False), [-3, -2])


# --------------------------------------------------------------------------
# a9-a11  prostate T2 chain (GRAPPA excluded -- SURVEY.md section 8a row a11)
# --------------------------------------------------------------------------


def ifftnd(kspace: np.ndarray, axes: Optional[Sequence[int]] = (-1,)) -> np.ndarray:
    """Centred inverse FFT scaled to orthonormal by ``* sqrt(prod(shape[axes]))``
    (ZIP!/fastmri_prostate/reconstruction/utils.py:7-29)."""
    if axes is None:
        axes = range(kspace.ndim)
    axes = list(axes)
    img = np.fft.fftshift(np.fft.ifftn(np.fft.ifftshift(kspace, axes=axes), axes=axes), axes=axes)
    img *= np.sqrt(np.prod(np.take(img.shape, axes)))
    return img


def flip_im(vol: np.ndarray, slice_axis: int) -> np.ndarray:
    """ZIP!/fastmri_prostate/reconstruction/utils.py:32-51, including its quirk: the loop runs
    ``vol.shape[slice_axis]`` times but always indexes the FIRST axis; in place, returns ``vol``."""
    for i in range(vol.shape[slice_axis]):
        vol[i] = np.flipud(vol[i])
    return vol


def complex_center_crop_ri(data: np.ndarray, shape: Tuple[int, int]) -> np.ndarray:
    """Centre crop of dims -3, -2 of a real-view array (..., H, W, 2) (transforms.py:70-92)."""
    if not (0 < shape[0] <= data.shape[-3] and 0 < shape[1] <= data.shape[-2]):
        raise ValueError("Invalid shapes.")
    r0 = crop_start(data.shape[-3], shape[0])
    c0 = crop_start(data.shape[-2], shape[1])
    return data[..., r0:r0 + shape[0], c0:c0 + shape[1], :]


def center_crop_to_smallest(x: np.ndarray, y: np.ndarray):
    """Crop both to the smaller extent of each of the last two dims (transforms.py:95-117)."""
    w = min(x.shape[-1], y.shape[-1])
    h = min(x.shape[-2], y.shape[-2])
    return center_crop(x, (h, w)), center_crop(y, (h, w))


def create_coil_combined_im(k: np.ndarray) -> np.ndarray:
    """(S, C, RO, PE) complex -> (S, RO, PE) float64: per slice ifftnd over
    (RO, PE), RSS over coils, ``np.flipud`` (prostate_t2_recon.py:80-102)."""
    out = np.zeros((k.shape[0], k.shape[2], k.shape[3]))
    for s in range(out.shape[0]):
        coil_imgs = ifftnd(k[s], [1, 2])
        out[s] = np.flipud(rss_np(coil_imgs, axis=0))
    return out


def padding_lr(enc_x: int, max_pe_index: int) -> Tuple[int, int]:
    """Phase-encode zero-pad rule from the ISMRMRD header numbers
    (ZIP!/fastmri_prostate/data/mri_data.py:63-85 ``get_padding`` and :150-157):
    ``p = (enc_x - (max_pe_index + 1)) / 2``; floor/ceil when ``p % 2 != 0``.
    451 -> 640 gives (94, 95)."""
    p = (enc_x - (max_pe_index + 1)) / 2
    if p % 2 != 0:
        return int(np.floor(p)), int(np.ceil(p))
    return int(p), int(p)


def zero_pad_pe(k: np.ndarray, left: int, right: int) -> np.ndarray:
    """np.pad on the last (PE) axis of a 4-D array (mri_data.py:158)."""
    return np.pad(k, ((0, 0), (0, 0), (0, 0), (left, right)))


def t2_average_combine(kspace: np.ndarray, pad: Tuple[int, int],
                       crop: Tuple[int, int] = (320, 320)) -> np.ndarray:
    """(A, S, C, RO, PE) -> (S, crop, crop) float64 (prostate_t2_recon.py:65-75):
    for each average pad -> coil-combine; mean over averages AFTER the RSS; crop."""
    n_avg, n_sl, _, n_ro, _ = kspace.shape
    im = np.zeros((n_avg, n_sl, n_ro, n_ro))
    for a in range(n_avg):
        im[a] = create_coil_combined_im(zero_pad_pe(kspace[a], pad[0], pad[1]))
    return center_crop_im(np.mean(im, axis=0), list(crop))


# --------------------------------------------------------------------------
# a12  undersampling mask -- NOT IN THE REFERENCE (parity unpinned for generation)
# --------------------------------------------------------------------------


def equispaced_mask(width: int, acceleration: int, center_fraction: float,
                    offset: int = 0) -> np.ndarray:
    """Builder-defined equispaced mask (SURVEY.md section 8c last row):
    every ``acceleration``-th column from ``offset`` plus a centred block of
    ``round(width * center_fraction)`` low-frequency columns starting at
    ``(width - n_low + 1) // 2``.  float32 0/1 vector of length ``width``."""
    m = np.zeros(width, dtype=np.float32)
    m[offset::acceleration] = 1.0
    n_low = int(round(width * center_fraction))
    lo = (width - n_low + 1) // 2
    m[lo:lo + n_low] = 1.0
    return m


def apply_mask(kspace: np.ndarray, mask_w: Optional[np.ndarray]) -> np.ndarray:
    """fastMRI convention: multiply by a vector broadcast along the last axis."""
    if mask_w is None:
        return kspace
    return kspace * mask_w.astype(np.float32).reshape((1,) * (kspace.ndim - 1) + (-1,))


# --------------------------------------------------------------------------
# Compositions the CUDA path is checked against (the three chains of BASELINE.md section 3)
# --------------------------------------------------------------------------


def knee_chain_numpy(kspace: np.ndarray, mask_w: Optional[np.ndarray],
                     crop: Tuple[int, int] = (320, 320), normalize_mode: Optional[str] = "instance",
                     eps: float = 0.0):
    """(…, C, H, W) c64 -> (…, oh, ow) f32 via the ``src/utils/kspace.py`` functions:
    mask -> ifft2c -> complex_abs -> sqrt(sum^2 over coils) -> center_crop_or_pad
    (-> normalize_instance).  Returns (image, mean, std); mean/std are None when
    ``normalize_mode`` is None."""
    img = complex_abs(ifft2c(apply_mask(kspace, mask_w)))
    comb = np.sqrt((img ** 2).sum(axis=-3))
    comb = center_crop_or_pad(comb, crop[0], crop[1]).astype(np.float32, copy=False)
    if normalize_mode is None:
        return comb, None, None
    if comb.ndim == 2:
        return normalize_instance(comb, eps)
    outs = [normalize_instance(c, eps) for c in comb.reshape((-1,) + comb.shape[-2:])]
    lead = comb.shape[:-2]
    return (np.stack([o[0] for o in outs]).reshape(comb.shape),
            np.array([o[1] for o in outs], dtype=np.float32).reshape(lead),
            np.array([o[2] for o in outs], dtype=np.float32).reshape(lead))


def knee_chain_fastmri(kspace: np.ndarray, mask_w: Optional[np.ndarray],
                       crop: Tuple[int, int] = (320, 320), eps: float = 0.0):
    """(C, H, W) c64 -> (oh, ow) f32 via the vendored fastMRI functions:
    to_tensor -> mask -> ifft2c_new -> rss_complex -> center_crop -> normalize_instance."""
    ri = to_real_view(apply_mask(kspace, mask_w)).astype(np.float32, copy=False)
    comb = rss_complex_ri(ifft2c_new(ri), dim=0)
    return normalize_instance(np.ascontiguousarray(center_crop(comb, crop)), eps)


def prostate_chain(kspace: np.ndarray, mask_w: Optional[np.ndarray], pad: Tuple[int, int],
                   crop: Tuple[int, int] = (320, 320)) -> np.ndarray:
    """(A, S, C, RO, PE) c64 -> (S, oh, ow) f64: mask -> :func:`t2_average_combine`."""
    return t2_average_combine(apply_mask(kspace, mask_w), pad, crop)


# --------------------------------------------------------------------------- after the reconstruction
# REF/src/preprocess/mri_preprocess.py:59-84: clip -> (body mask) -> resize -> z-score -> preview

def percentile_clip(img: np.ndarray, pmin: float, pmax: float):
    """``_percentile_clip`` (``mri_preprocess.py:182-185``); also returns (lo, hi)."""
    lo, hi = np.percentile(img, pmin), np.percentile(img, pmax)
    return np.clip(img, lo, hi), lo, hi


def percentile_f32(img: np.ndarray, q: float) -> np.float32:
    """What ``np.percentile`` (numpy >= 2.0, method 'linear') does to a float32 array, spelled out: the quantile, the
    virtual index and the interpolation are all float32 -- the arithmetic the CUDA kernel restates."""
    f32 = np.float32
    s = np.sort(np.asarray(img, dtype=f32).ravel())
    n = s.size
    v = f32(f32(n - 1) * (f32(q) / f32(100)))
    i = int(np.floor(v))
    g = f32(v - f32(i))
    a, b = s[min(i, n - 1)], s[min(i + 1, n - 1)]
    d = f32(b - a)
    return f32(a + f32(d * g)) if g < f32(0.5) else f32(b - f32(d * f32(f32(1) - g)))


def resize_bilinear(img: np.ndarray, out_hw: Tuple[int, int]) -> np.ndarray:
    """``_resize_np`` (``mri_preprocess.py:187-191``): ``F.interpolate(mode='bilinear', align_corners=False)`` restated in
    numpy float32 (ATen ``area_pixel_compute_source_index``: scale = in / out, s = scale * (d + 0.5) - 0.5 clamped at
    0; taps i0 = int(s), i1 = i0 + (i0 < in - 1); value = h0 * (w0 p00 + w1 p01) + h1 * (w0 p10 + w1 p11))."""
    f32 = np.float32
    x = np.asarray(img, dtype=f32)
    H, W = x.shape
    oh, ow = int(out_hw[0]), int(out_hw[1])

    def taps(n_in, n_out):
        scale = f32(n_in) / f32(n_out)
        s = scale * (np.arange(n_out, dtype=f32) + f32(0.5)) - f32(0.5)
        s = np.where(s < 0, f32(0), s).astype(f32)
        i0 = np.minimum(s.astype(np.int64), n_in - 1)
        i1 = i0 + (i0 < n_in - 1)
        l1 = (s - i0.astype(f32)).astype(f32)
        return i0, i1, (f32(1) - l1).astype(f32), l1

    y0, y1, hy, ly = taps(H, oh)
    x0, x1, hx, lx = taps(W, ow)
    top = hx[None, :] * x[y0][:, x0] + lx[None, :] * x[y0][:, x1]
    bot = hx[None, :] * x[y1][:, x0] + lx[None, :] * x[y1][:, x1]
    return (hy[:, None] * top + ly[:, None] * bot).astype(f32)


def resize_mask(mask: np.ndarray, out_hw: Tuple[int, int]) -> np.ndarray:
    """``(self._resize_np(mk.astype(np.float32), self.out_size) > 0.5).astype(np.uint8)`` (``mri_preprocess.py:77``)."""
    return (resize_bilinear(mask.astype(np.float32), out_hw) > 0.5).astype(np.uint8)


def zscore_in_mask(img: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """``_zscore_in_mask`` (``mri_preprocess.py:216-224``): population std, < 10 pixels -> whole image, std floor."""
    vals = img[mask > 0]
    if vals.size < 10:
        mean, std = img.mean(), img.std()
    else:
        mean, std = vals.mean(), vals.std()
    std = std if std > 1e-6 else 1.0
    return ((img - mean) / std).astype(np.float32)


def preview_01(img: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """``_preview_01`` (``mri_preprocess.py:226-233``)."""
    vals = img[mask > 0]
    if vals.size > 0:
        lo, hi = float(vals.min()), float(vals.max())
    else:
        lo, hi = float(img.min()), float(img.max())
    return ((img - lo) / (hi - lo + 1e-6)).astype(np.float32)


def post_chain(img: np.ndarray, body_mask: Optional[np.ndarray], out_hw: Tuple[int, int] = (320, 320),
               clip: Tuple[float, float] = (1.0, 99.5)):
    """The tail of ``preprocess_record`` (``mri_preprocess.py:62-84``) with the body mask given (Otsu / morphology need
    scikit-image): clip -> resize image and mask -> z-score -> preview.  Returns (img_z, img_01, mask_r, (lo, hi))."""
    clipped, lo, hi = percentile_clip(img, *clip)
    img_r = resize_bilinear(clipped, out_hw)
    mk = np.ones(img.shape, np.uint8) if body_mask is None else body_mask
    mk_r = resize_mask(mk, out_hw)
    return zscore_in_mask(img_r, mk_r), preview_01(img_r, mk_r), mk_r, (np.float32(lo), np.float32(hi))


# --------------------------------------------------------------------------- GRAPPA / SENSE
def grappa_geometries(kspace: np.ndarray, kernel_size: Tuple[int, int] = (5, 5), coil_axis: int = -1):
    """``Grappa.get_kernel_geometries`` (``ZIP!/fastmri_prostate/reconstruction/grappa.py:15-102``) restated literally:
    pad, mask = |coil 0| > 0, all kx x ky windows, ``np.unique(axis=0)``, keep geometries with a hole at the centre that
    are not empty.  Returns (padded k-space (X+2kx2, Y+2ky2, C), patches (n, kx, ky) bool, patch_indices, holes_x, holes_y)."""
    k = np.moveaxis(kspace, coil_axis, -1)
    kx, ky = kernel_size
    kx2, ky2 = int(kx / 2), int(ky / 2)
    k = np.pad(k, ((kx2, kx2), (ky2, ky2), (0, 0)), mode="constant")
    mask = np.ascontiguousarray(np.abs(k[..., 0]) > 0)
    P = np.lib.stride_tricks.sliding_window_view(mask, (kx, ky))
    psh = P.shape[:2]
    P, iidx = np.unique(P.reshape((-1, kx, ky)), return_inverse=True, axis=0)
    iidx = np.asarray(iidx).reshape(-1)
    valid = np.argwhere(~P[:, kx2, ky2]).squeeze()
    invalid = np.argwhere(np.all(P == 0, axis=(1, 2)))
    valid = np.atleast_1d(np.setdiff1d(valid, invalid, assume_unique=True))
    hx, hy = {}, {}
    for ii in valid:
        x, y = np.unravel_index(np.flatnonzero(iidx == ii), psh)
        hx[ii], hy[ii] = x + kx2, y + ky2
    return k, P, valid, hx, hy


def grappa_weights(calib: np.ndarray, P: np.ndarray, patch_indices, kernel_size=(5, 5), coil_axis: int = -1, lamda: float = 0.01):
    """``Grappa.compute_weights`` (``grappa.py:104-171``)."""
    calib = np.moveaxis(calib, coil_axis, -1)
    kx, ky = kernel_size
    kx2, ky2 = int(kx / 2), int(ky / 2)
    nc = calib.shape[-1]
    calib = np.pad(calib, ((kx2, kx2), (ky2, ky2), (0, 0)), mode="constant")
    A = np.lib.stride_tricks.sliding_window_view(calib, (kx, ky, nc)).reshape((-1, kx, ky, nc))
    Pc = np.tile(P[..., None], (1, 1, 1, nc))
    out = {}
    for ii in patch_indices:
        S = A[:, Pc[ii, ...]]
        T = A[:, kx2, ky2, :]
        ShS = S.conj().T @ S
        ShT = S.conj().T @ T
        lamda0 = lamda * np.linalg.norm(ShS) / ShS.shape[0]
        out[ii] = np.linalg.solve(ShS + lamda0 * np.eye(ShS.shape[0]), ShT).T
    return out


def grappa_apply(kspace: np.ndarray, weights, kernel_size=(5, 5), coil_axis: int = -1) -> np.ndarray:
    """``Grappa.apply_weights`` (``grappa.py:173-222``) for the geometries of ``kspace`` itself: every hole of geometry ii gets
    ``weights[ii] @ S`` (S = sampled window entries, window-position major / coil minor); result = recon + kspace, in the
    input's dtype.  Vectorised over the holes of one geometry; same arithmetic (complex128 weights times complex64 sources)."""
    kx, ky = kernel_size
    kx2, ky2 = int(kx / 2), int(ky / 2)
    k, P, valid, hx, hy = grappa_geometries(kspace, kernel_size, coil_axis)
    recon = np.zeros(k.shape, dtype=k.dtype)
    for ii in valid:
        pi, pj = np.nonzero(P[ii])
        xs, ys = hx[ii], hy[ii]
        S = k[xs[:, None] + pi[None, :] - kx2, ys[:, None] + pj[None, :] - ky2, :]        # (holes, n_s, C)
        recon[xs, ys, :] = S.reshape(len(xs), -1) @ np.asarray(weights[ii]).T
    return np.moveaxis((recon + k)[kx2:-kx2, ky2:-ky2, :], -1, coil_axis)


def t2_reconstruction(kspace_data: np.ndarray, calib_data: np.ndarray, pad: Tuple[int, int],
                      crop: Tuple[int, int] = (320, 320)) -> np.ndarray:
    """``t2_reconstruction`` (``ZIP!/fastmri_prostate/reconstruction/t2/prostate_t2_recon.py:9-78``) restated with the
    functions above: geometries from slice 0 of averages 0 and 1, per-slice weights for both, averages 0/1/2 filled with
    objects 1/2/1 into a complex128 array, then :func:`t2_average_combine`.  Returns ``(S, oh, ow)`` float64."""
    num_avg, num_slices = kspace_data.shape[:2]
    geo = []
    for a in (0, 1):
        _, P, valid, _, _ = grappa_geometries(np.transpose(kspace_data[a, 0], (2, 0, 1)), (5, 5), 1)
        geo.append((P, valid))
    weights = [[grappa_weights(np.transpose(calib_data[s], (2, 0, 1)), P, valid, (5, 5), 1) for s in range(num_slices)]
               for P, valid in geo]
    post = np.zeros(kspace_data.shape, dtype=complex)
    for average, gi in zip((0, 1, 2), (0, 1, 0)):
        for s in range(num_slices):
            filled = grappa_apply(np.transpose(kspace_data[average, s], (2, 0, 1)), weights[gi][s], (5, 5), 1)
            post[average, s] = np.moveaxis(np.moveaxis(filled, 0, 1), 1, 2)
    return t2_average_combine(post, pad, crop)


def sense_combine(img: np.ndarray, sens: np.ndarray, magnitude: bool = True) -> np.ndarray:
    """``np.sum(img * sens.conj(), axis=1)`` (+ ``np.abs``): ``ZIP!/fastmri_prostate/reconstruction/dwi/prostate_dwi_recon.py:106-109``;
    ``sens_reduce`` of ``ZIP!/DL_reconstruction/models/varnet.py:199-203`` is the same sum on real views (coil axis 1)."""
    out = np.sum(img * sens.conj(), axis=1)
    return np.abs(out) if magnitude else out


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    """||a-b||_2 / ||b||_2 in float64 -- the parity metric of BASELINE.json."""
    a = np.asarray(a)
    b = np.asarray(b)
    wide = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    d = (a.astype(wide) - b.astype(wide)).ravel()
    den = math.sqrt(float(np.vdot(b.astype(wide).ravel(), b.astype(wide).ravel()).real))
    return math.sqrt(float(np.vdot(d, d).real)) / (den if den > 0 else 1.0)
