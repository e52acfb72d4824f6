"""Seeded synthetic k-space and sampling masks.

Host-side input generation for tests, ``bench.py`` and the golden-vector script.
Not on the hot path: nothing here runs inside a timed region.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

KNEE_SHAPE = (15, 640, 368)          # coils, readout, phase-encode  (BASELINE.json configs[0..1])
PROSTATE_SHAPE = (3, 30, 16, 640, 451)  # averages, slices, coils, readout, phase-encode (configs[2])
PROSTATE_PAD = (94, 95)              # 451 -> 640
CROP = (320, 320)


def gaussian_kspace(shape: Tuple[int, ...], seed: int) -> np.ndarray:
    """complex64 ``N(0,1) + i N(0,1)`` of ``shape`` from ``default_rng(seed)``."""
    rng = np.random.default_rng(seed)
    re = rng.standard_normal(shape, dtype=np.float32)
    im = rng.standard_normal(shape, dtype=np.float32)
    out = np.empty(shape, dtype=np.complex64)
    out.real = re
    out.imag = im
    return out


def phantom_kspace(shape: Tuple[int, int, int], seed: int) -> np.ndarray:
    """Structured (C, H, W) case: centred orthonormal FFT of an ellipse phantom times
    smooth Gaussian coil sensitivities, so the RSS image has image-like dynamic range."""
    c, h, w = shape
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(-1, 1, h, dtype=np.float32),
                         np.linspace(-1, 1, w, dtype=np.float32), indexing="ij")
    body = ((xx / 0.62) ** 2 + (yy / 0.42) ** 2 <= 1).astype(np.float32)
    body += 0.6 * (((xx + 0.2) / 0.2) ** 2 + ((yy - 0.05) / 0.12) ** 2 <= 1)
    body -= 0.4 * (((xx - 0.25) / 0.15) ** 2 + ((yy + 0.1) / 0.1) ** 2 <= 1)
    body *= 1.0 + 0.05 * np.sin(9 * xx) * np.cos(7 * yy)
    ang = 2 * np.pi * np.arange(c) / c + rng.uniform(0, 0.3)
    imgs = np.empty((c, h, w), dtype=np.complex64)
    for i in range(c):
        cx, cy = 0.9 * np.cos(ang[i]), 0.9 * np.sin(ang[i])
        sens = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / 0.9).astype(np.float32)
        phase = np.exp(1j * (0.8 * xx * np.cos(ang[i]) + 0.8 * yy * np.sin(ang[i]))).astype(np.complex64)
        imgs[i] = body * sens * phase
    ax = (-2, -1)
    k = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(imgs, axes=ax), norm="ortho"), axes=ax)
    noise = 1e-3 * gaussian_kspace((c, h, w), seed + 7919)
    return (k + noise).astype(np.complex64)


def equispaced_mask(width: int, acceleration: int, center_fraction: float, offset: int = 0) -> np.ndarray:
    """0/1 float32 mask along the phase-encode axis.

    The reference has no mask code; this is the rule
    this build defines and freezes: every ``acceleration``-th column from
    ``offset``, plus ``round(width * center_fraction)`` centred low-frequency columns
    starting at ``(width - n_low + 1) // 2``.  (368, 4, 0.08) keeps 114 columns."""
    m = np.zeros(width, dtype=np.float32)
    m[offset::acceleration] = 1.0
    n_low = int(round(width * center_fraction))
    lo = (width - n_low + 1) // 2
    m[lo:lo + n_low] = 1.0
    return m


def knee_mask() -> np.ndarray:
    """configs[0..1]: W=368, 4x equispaced + 29 ACS columns (114 kept)."""
    return equispaced_mask(368, 4, 0.08)


def prostate_mask() -> np.ndarray:
    """configs[2]: PE=451, 8x equispaced + round(451*0.04)=18 ACS columns."""
    return equispaced_mask(451, 8, 0.04)


def prostate_volume_block(a: int, s: int, seed: int = 0) -> np.ndarray:
    """``(C, RO, PE)`` block (average ``a``, slice ``s``) of the configs[2] volume ``PROSTATE_SHAPE``.

    The full-size parity tests and the golden script build the 3.3 GB volume block by block from these
    seeds, so neither needs the whole volume in host memory at once."""
    _, n_sl, c, ro, pe = PROSTATE_SHAPE
    return gaussian_kspace((c, ro, pe), 100000 + 1000 * seed + a * n_sl + s)


#: (width, acceleration, center_fraction, offset) of every sampling mask the tests and the bench use; their index
#: lists are frozen in tests/golden/manifest.json ("masks") -- the generator is builder-defined
FROZEN_MASKS = ((368, 4, 0.08, 0), (368, 8, 0.04, 0), (368, 4, 0.08, 1), (368, 4, 0.08, 3), (640, 8, 0.04, 0),
                (451, 8, 0.04, 0), (372, 4, 0.08, 0))


def mask_name(width: int, acceleration: int, center_fraction: float, offset: int = 0) -> str:
    return f"equispaced({width},{acceleration},{center_fraction:g},{offset})"


def magnitude_image(shape: Tuple[int, int], seed: int, flat: bool = False) -> np.ndarray:
    """float32 ``(H, W)`` magnitude-like image for the steps after the reconstruction: an off-centre ellipse with smooth
    shading and Rician-looking noise (``flat``: a constant image, the std-floor case).  Pure numpy from ``seed``."""
    h, w = shape
    if flat:
        return np.full(shape, 0.25, dtype=np.float32)
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(-1, 1, h, dtype=np.float32), np.linspace(-1, 1, w, dtype=np.float32), indexing="ij")
    body = (((xx - 0.1) / 0.7) ** 2 + ((yy + 0.05) / 0.55) ** 2 <= 1).astype(np.float32)
    body += 0.5 * (((xx + 0.2) / 0.2) ** 2 + ((yy - 0.1) / 0.15) ** 2 <= 1)
    shade = (1.0 + 0.3 * xx - 0.2 * yy).astype(np.float32)
    n1 = rng.standard_normal(shape, dtype=np.float32)
    n2 = rng.standard_normal(shape, dtype=np.float32)
    return np.sqrt((body * shade + 0.03 * n1) ** 2 + (0.03 * n2) ** 2).astype(np.float32)


def body_mask_standin(img: np.ndarray, frac: float = 0.3) -> np.ndarray:
    """uint8 mask ``img > frac * max``: a stand-in for the reference's Otsu + morphology body mask (scikit-image is not
    available here; the steps under test only need SOME mask)."""
    return (img > frac * img.max()).astype(np.uint8)


#: (name, input shape, output size, seed, mask rule) of the post-reconstruction golden cases (tests/golden/post_vectors.npz)
POST_CASES = (("knee_640x368", (640, 368), (320, 320), 501, "standin"),
              ("knee_320x320_identity", (320, 320), (320, 320), 502, "standin"),
              ("odd_37x53_to_24x40", (37, 53), (24, 40), 503, "standin"),
              ("tiny_mask", (37, 53), (24, 40), 504, "tiny"),          # < 10 pixels inside: whole-image statistics
              ("empty_mask", (37, 53), (24, 40), 505, "empty"),        # preview falls back to the image extrema
              ("flat_image", (32, 32), (16, 16), 0, "flat"))           # std <= 1e-6 -> 1


def post_case_inputs(name: str):
    for nm, shape, out, seed, rule in POST_CASES:
        if nm == name:
            img = magnitude_image(shape, seed, flat=(rule == "flat"))
            if rule in ("standin", "flat"):
                mk = body_mask_standin(img) if rule == "standin" else np.ones(shape, np.uint8)
            elif rule == "tiny":
                mk = np.zeros(shape, np.uint8)
                mk[10:13, 20:23] = 1
            else:
                mk = np.zeros(shape, np.uint8)
            return img, mk, out
    raise KeyError(name)


#: GRAPPA golden cases: (name, PE, coils, RO, acceleration, ACS lines, calibration PE lines, seed); data layout (PE, coil, RO)
#: with coil_axis = 1, the layout the T2 / DWI chains hand to Grappa (prostate_t2_recon.py:34)
GRAPPA_CASES = (("small_r2", 24, 4, 20, 2, 6, 12, 601), ("medium_r3", 90, 8, 64, 3, 12, 24, 602))


def grappa_case_inputs(name: str):
    """(undersampled k-space (PE, C, RO) c64 with exact zeros on the skipped lines, calibration (PE_cal, C, RO) c64)."""
    for nm, pe, nc, ro, acc, acs, cal, seed in GRAPPA_CASES:
        if nm == name:
            k = gaussian_kspace((pe, nc, ro), seed)
            keep = np.zeros(pe, dtype=bool)
            keep[::acc] = True
            lo = (pe - acs) // 2
            keep[lo:lo + acs] = True
            k[~keep] = 0
            # calibration with structure (a smooth kernel correlates neighbours, so the fit is well conditioned)
            c = gaussian_kspace((cal + 4, nc, ro + 4), seed + 1)
            calib = (c[2:-2, :, 2:-2] + 0.5 * (c[1:-3, :, 2:-2] + c[3:-1, :, 2:-2]) + 0.5 * (c[2:-2, :, 1:-3] + c[2:-2, :, 3:-1])).astype(np.complex64)
            return k, calib
    raise KeyError(name)


#: full T2 reconstruction case (prostate_t2_recon.py:9-78): (averages, slices, coils, RO, PE), ACS lines, calibration lines,
#: seed.  RO = 320 so that the reference's hard-wired (320, 320) crop is the whole image; PE 60 -> padding 130 | 130.
#: PE is kept small on purpose: numpy 2.3.5's np.unravel_index returns wrong coordinates for an (m, 1)-shaped index array
#: with m > 8192, which is what Grappa.get_kernel_geometries (grappa.py:88-90) hands it -- under this numpy the vendored
#: class silently leaves every hole after the 8192nd of a geometry unfilled (tests/test_grappa_sense.py shows it).  The
#: golden cases stay below that size so that they pin the algorithm, not the library bug.
T2_RECON_CASE = ((3, 2, 4, 320, 60), 12, 24, 701)


def t2_recon_case_inputs():
    """(kspace (3,S,C,RO,PE) c64 with R = 2 interleaved between odd / even averages + ACS, calibration (S,C,RO,PE_cal), header)."""
    (na, ns, nc, ro, pe), acs, cal, seed = T2_RECON_CASE
    k = gaussian_kspace((na, ns, nc, ro, pe), seed)
    lo = (pe - acs) // 2
    for a in range(na):
        keep = np.zeros(pe, dtype=bool)
        keep[(a % 2)::2] = True               # the second average samples the other set of lines
        keep[lo:lo + acs] = True
        k[a][..., ~keep] = 0
    c = gaussian_kspace((ns, nc, ro + 4, cal + 4), seed + 1)
    calib = (c[:, :, 2:-2, 2:-2] + 0.5 * (c[:, :, 1:-3, 2:-2] + c[:, :, 3:-1, 2:-2]) + 0.5 * (c[:, :, 2:-2, 1:-3] + c[:, :, 2:-2, 3:-1])).astype(np.complex64)
    hdr = ("<?xml version=\"1.0\"?><ismrmrdHeader xmlns=\"http://www.ismrm.org/ISMRMRD\"><encoding><encodedSpace><matrixSize>"
           f"<x>{ro}</x><y>{pe}</y><z>1</z></matrixSize></encodedSpace><encodingLimits><kspace_encoding_step_1><minimum>0</minimum>"
           f"<maximum>{pe - 1}</maximum><center>{pe // 2}</center></kspace_encoding_step_1></encodingLimits></encoding></ismrmrdHeader>")
    return k, calib, hdr
