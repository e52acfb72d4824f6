"""Parity of the sm_100a library against the oracle and the frozen reference outputs, through
the public Python API (which calls the C ABI).  Run on the B200 box: pytest -m gpu.

Bar (BASELINE.json): rel-L2 <= 1e-5 for floating point; mask and crop indexing bit-exact.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.adapters import recon_cabi as cabi
from mri_acl_imagesegmentation_adsp_b200.fastmri import coil_combine, fftc, math_fn, transforms
from mri_acl_imagesegmentation_adsp_b200.preprocess.mri_preprocess import MRIKneePreprocessor
from mri_acl_imagesegmentation_adsp_b200.prostate import t2
from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import recon_to_unet_input, zero_filled_rss
from mri_acl_imagesegmentation_adsp_b200.utils import kspace as K
from oracle import recon_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _experimental_built() -> bool:
    """the schedules that were measured slower than the default live behind -DMRIACL_EXPERIMENTAL
    (`make -C csrc experimental`, MRIACL_RECON_LIBRARY=...); the product library answers "schedule not built"."""
    k = torch.zeros((1, 1, 640, 368), dtype=torch.complex64, device="cuda")
    try:
        zero_filled_rss(k, None, synth.CROP, None, schedule="fused")
        return True
    except ValueError as e:
        assert "not built" in str(e)
        return False


experimental = pytest.mark.skipif("not _experimental_built()", reason="library built without MRIACL_EXPERIMENTAL")


@pytest.fixture(scope="module", autouse=True)
def _native_library_loaded():
    lib = cabi.library()          # raises if csrc/libmriacl_recon.so is missing: no silent fallback
    before = lib.launch_count()
    yield
    assert lib.launch_count() > before, "no kernel of libmriacl_recon.so was launched"


# ---------------------------------------------------------------------------------------------
# complex-output API (generic kernels)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(4, 32, 24), (2, 3, 30, 23), (1, 37, 5), (2, 640, 368), (1, 640, 372), (3, 451)])
def test_fft2c_ifft2c(shape):
    x = synth.gaussian_kspace(shape, 31)
    a, b = K.ifft2c(x), K.fft2c(x)
    assert isinstance(a, np.ndarray) and a.dtype == np.complex64 and a.shape == x.shape
    assert O.rel_l2(a, O.ifft2c(x)) <= TOL
    assert O.rel_l2(b, O.fft2c(x)) <= TOL
    assert O.rel_l2(K.fft2c(K.ifft2c(x)), x) <= TOL


def test_golden_small(golden):
    k = golden["small_even/kspace"]
    assert O.rel_l2(K.ifft2c(k), golden["small_even/ifft2c"]) <= TOL
    assert O.rel_l2(K.fft2c(k), golden["small_even/fft2c"]) <= TOL
    assert O.rel_l2(K.complex_abs(k), golden["small_even/complex_abs"]) <= 1e-6
    ko = golden["small_odd/kspace"]
    assert O.rel_l2(K.ifft2c(ko), golden["small_odd/ifft2c"]) <= TOL
    t = transforms.to_tensor(ko)
    assert O.rel_l2(fftc.ifft2c_new(t).numpy(), golden["small_odd/ifft2c_new"]) <= TOL
    assert O.rel_l2(fftc.fft2c_new(t).numpy(), golden["small_odd/fft2c_new"]) <= TOL
    img = fftc.ifft2c_new(t.cuda())
    assert img.is_cuda and img.shape == t.shape
    assert O.rel_l2(coil_combine.rss_complex(img, dim=1).cpu().numpy(), golden["small_odd/rss_complex"]) <= TOL
    assert O.rel_l2(coil_combine.rss(torch.from_numpy(np.abs(ko)), dim=1).numpy(), golden["small_odd/rss_real"]) <= 1e-6
    assert O.rel_l2(math_fn.complex_abs(t).numpy(), np.abs(ko)) <= 1e-6
    assert O.rel_l2(math_fn.complex_abs_sq(t).numpy(), np.abs(ko) ** 2) <= 1e-6
    assert O.rel_l2(t2.ifftnd(ko[0], [1, 2]), golden["small_odd/ifftnd"]) <= TOL
    assert O.rel_l2(MRIKneePreprocessor.ifft2c_single(ko[1, 2]), golden["small_odd/ifft2c_single"]) <= TOL
    with pytest.raises(ValueError):
        fftc.ifft2c_new(torch.zeros(4, 4, 3))
    with pytest.raises(ValueError):
        MRIKneePreprocessor.ifft2c_single(ko)


def test_crop_pad_bit_exact(golden):
    x = golden["small_even/ifft2c_single_coil0"]
    np.testing.assert_array_equal(K.center_crop_or_pad(x, 40, 12), golden["small_even/crop_or_pad_40x12"])
    probe = np.arange(640 * 368, dtype=np.float32).reshape(640, 368)
    c = K.center_crop_or_pad(probe, 320, 320)
    np.testing.assert_array_equal(c, probe[160:480, 24:344])
    np.testing.assert_array_equal(c[[0, 0, -1, -1], [0, -1, 0, -1]], golden["index/crop_probe_corners"])
    z = synth.gaussian_kspace((2, 3, 30, 23), 4)
    for oh, ow in [(12, 40), (30, 23), (31, 22), (7, 7)]:
        np.testing.assert_array_equal(K.center_crop_or_pad(z, oh, ow), O.center_crop_or_pad(z, oh, ow))
    np.testing.assert_array_equal(transforms.center_crop(torch.from_numpy(probe), (320, 320)).numpy(), c)
    with pytest.raises(ValueError):
        transforms.center_crop(torch.from_numpy(probe), (641, 320))


def test_normalize_instance(golden):
    x = np.abs(synth.gaussian_kspace((320, 320), 3)).astype(np.float32) + 2.0
    out, mean, std = transforms.normalize_instance(torch.from_numpy(x), eps=1e-11)
    np.testing.assert_allclose([mean.item(), std.item()], golden["norm/mean_std"], rtol=2e-6)
    assert O.rel_l2(out.numpy()[::8, ::8], golden["norm/out_sub"]) <= TOL


# ---------------------------------------------------------------------------------------------
# the fused stage, configs[0]
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,gen", [("knee_gauss", synth.gaussian_kspace), ("knee_phantom", synth.phantom_kspace)])
def test_knee_config0_against_reference_outputs(golden, name, gen):
    k = gen(synth.KNEE_SHAPE, 0)
    m = synth.knee_mask()
    img, mean, std = zero_filled_rss(k, m, synth.CROP, "instance")
    assert img.dtype == np.float32 and img.shape == (320, 320)
    assert O.rel_l2(img, golden[f"{name}/fastmri_chain"]) <= TOL
    np.testing.assert_allclose([mean, std], golden[f"{name}/fastmri_mean_std"], rtol=1e-5)
    raw, _, _ = zero_filled_rss(k, m, synth.CROP, None)
    assert O.rel_l2(raw, golden[f"{name}/numpy_chain"]) <= TOL
    gen_raw, _, _ = zero_filled_rss(k, m, synth.CROP, None, force_generic=True)
    assert O.rel_l2(gen_raw, golden[f"{name}/numpy_chain"]) <= TOL
    dense, _, _ = zero_filled_rss(k, None, synth.CROP, None)
    if name == "knee_gauss":
        assert O.rel_l2(dense[::4, ::4], golden["knee_gauss/numpy_chain_nomask_sub4"]) <= TOL


def test_mask_and_crop_indexing_bit_exact():
    k = torch.from_numpy(synth.gaussian_kspace((2,) + synth.KNEE_SHAPE, 40)).cuda()
    m = synth.knee_mask()
    raw, _, _ = zero_filled_rss(k, m, synth.CROP, None)
    full, _, _ = zero_filled_rss(k, m, None, None)
    assert full.shape == (2, 640, 368)
    assert torch.equal(raw, full[:, 160:480, 24:344])
    k2 = k.clone()
    k2[..., torch.from_numpy(m == 0).cuda()] = complex(1e30, -1e30)
    raw2, _, _ = zero_filled_rss(k2, m, synth.CROP, None)
    assert torch.equal(raw, raw2)
    # batching / chunking never changes a slice
    one, _, _ = zero_filled_rss(k[1], m, synth.CROP, None)
    assert torch.equal(one, raw[1])
    c1, _, _ = zero_filled_rss(k, m, synth.CROP, None, chunk_slices=1)
    assert torch.equal(c1, raw)
    # the experimental schedules (when built) and the back-to-back kernel schedule do the same arithmetic
    for schedule in ("sequential",) + (("fused", "overlapped") if _experimental_built() else ()):
        alt, _, _ = zero_filled_rss(k, m, synth.CROP, None, schedule=schedule)
        assert torch.equal(alt, raw), schedule
    with pytest.raises(ValueError):
        zero_filled_rss(k, m, synth.CROP, None, schedule="bogus")


@experimental
@pytest.mark.parametrize("schedule", ["fused", "overlapped"])
def test_experimental_schedules_many_groups(schedule):
    """the persistent-kernel schedules (dynamic work claiming, per-slice counters, side stream) give the same
    images as the back-to-back one, also when the batch is larger than the workspace (two alternating buffers)."""
    g = torch.Generator(device="cuda").manual_seed(3)
    k = torch.view_as_complex(torch.randn((13, 15, 640, 368, 2), device="cuda", generator=g))
    m = synth.knee_mask()
    ref, _, _ = zero_filled_rss(k, m, synth.CROP, None, schedule="sequential", chunk_slices=13)
    nref, rmean, rstd = zero_filled_rss(k, m, synth.CROP, "instance", schedule="sequential", chunk_slices=13)
    for chunk in (13, 6, 4, 2, 1):
        out, _, _ = zero_filled_rss(k, m, synth.CROP, None, chunk_slices=chunk, schedule=schedule)
        assert torch.equal(out, ref), chunk            # same arithmetic, bit for bit
        nout, mean, std = zero_filled_rss(k, m, synth.CROP, "instance", chunk_slices=chunk, schedule=schedule)
        # (tile statistics are reduced over a different number of warps / tiles: last-bit differences only)
        torch.testing.assert_close(nout, nref, rtol=0, atol=2e-6)
        torch.testing.assert_close(mean, rmean, rtol=1e-6, atol=0)
        torch.testing.assert_close(std, rstd, rtol=1e-6, atol=0)
    for _ in range(5):      # repeated calls reuse counters, events and the side stream
        out, _, _ = zero_filled_rss(k, m, synth.CROP, None, chunk_slices=5, schedule=schedule)
    assert torch.equal(out, ref)
    with pytest.raises(ValueError):
        zero_filled_rss(k, m, synth.CROP, None, schedule="bogus")


@experimental
def test_coresident_schedule():
    """the one-launch co-resident schedule (column team + row team in every CTA, fused normalisation) gives the same
    images as the back-to-back kernels: bit-identical before normalisation, last-bit statistics differences after;
    repeated calls and chunked workspaces reuse the counters correctly."""
    g = torch.Generator(device="cuda").manual_seed(5)
    k = torch.view_as_complex(torch.randn((13, 15, 640, 368, 2), device="cuda", generator=g))
    m = synth.knee_mask()
    ref, _, _ = zero_filled_rss(k, m, synth.CROP, None, schedule="sequential")
    nref, rmean, rstd = zero_filled_rss(k, m, synth.CROP, "instance", schedule="sequential")
    first = None
    for chunk in (13, 5, 1):
        out, _, _ = zero_filled_rss(k, m, synth.CROP, None, chunk_slices=chunk, schedule="coresident")
        assert O.rel_l2(out.cpu().numpy(), ref.cpu().numpy()) <= 2e-6, chunk
        first = out if first is None else first
        assert torch.equal(out, first), chunk          # chunking never changes a slice
        nout, mean, std = zero_filled_rss(k, m, synth.CROP, "instance", chunk_slices=chunk, schedule="coresident")
        torch.testing.assert_close(nout, nref, rtol=0, atol=2e-5)
        torch.testing.assert_close(mean, rmean, rtol=1e-5, atol=0)
        torch.testing.assert_close(std, rstd, rtol=1e-5, atol=0)
    for _ in range(4):
        out, _, _ = zero_filled_rss(k, m, synth.CROP, None, schedule="coresident")
    assert torch.equal(out, first)
    # (single average + pair-family mask above runs the pair row team; MRIACL_KC_PAIR=0 selects the cooperative one)
    # averages + flip + odd crop through the same kernel
    k5 = torch.view_as_complex(torch.randn((3, 2, 4, 640, 368, 2), device="cuda", generator=g))
    a, _, _ = zero_filled_rss(k5, m, (77, 200), None, average_axis=1, flip_rows=True, schedule="coresident")
    b, _, _ = zero_filled_rss(k5, m, (77, 200), None, average_axis=1, flip_rows=True, schedule="sequential")
    assert torch.equal(a, b)


@experimental
def test_pipelined_schedule():
    """small chunks on two streams with T double-buffered in L2: same bits as one big chunk, incl. the fused normalisation;
    ragged last chunk; repeated calls (event / buffer reuse)."""
    g = torch.Generator(device="cuda").manual_seed(9)
    k = torch.view_as_complex(torch.randn((21, 15, 640, 368, 2), device="cuda", generator=g))
    m = synth.knee_mask()
    ref, _, _ = zero_filled_rss(k, m, synth.CROP, None, schedule="sequential")
    nref, rmean, rstd = zero_filled_rss(k, m, synth.CROP, "instance", schedule="sequential")
    for chunk in (21, 16, 5):
        out, _, _ = zero_filled_rss(k, m, synth.CROP, None, chunk_slices=chunk, schedule="pipelined")
        assert torch.equal(out, ref), chunk
        nout, mean, std = zero_filled_rss(k, m, synth.CROP, "instance", chunk_slices=chunk, schedule="pipelined")
        assert torch.equal(nout, nref) and torch.equal(mean, rmean) and torch.equal(std, rstd)
    for _ in range(3):
        nout, _, _ = zero_filled_rss(k, m, synth.CROP, "instance", schedule="pipelined")
    assert torch.equal(nout, nref)
    want = O.knee_chain_numpy(k[20].cpu().numpy(), m, synth.CROP, "instance")
    assert O.rel_l2(nout[20].cpu().numpy(), want[0]) <= TOL


@experimental
def test_pair_row_pass_schedule():
    """the pair row pass (rowpair.cuh) against the oracle (rel-L2 <= 1e-5) and the cooperative row pass; bit-stable
    under chunking; masked columns never read; a mask offset that needs the index rotation; fallback outside its family."""
    k_np = synth.gaussian_kspace((5,) + synth.KNEE_SHAPE, 51)
    k = torch.from_numpy(k_np).cuda()
    m = synth.knee_mask()
    ref, _, _ = zero_filled_rss(k, m, synth.CROP, None, schedule="sequential")
    out, _, _ = zero_filled_rss(k, m, synth.CROP, None, schedule="pair")
    assert O.rel_l2(out.cpu().numpy(), ref.cpu().numpy()) <= 2e-6
    want = O.knee_chain_numpy(k_np[3], m, synth.CROP, None)
    want = want[0] if isinstance(want, tuple) else want
    assert O.rel_l2(out[3].cpu().numpy(), want) <= TOL
    for chunk in (1, 2, 5):
        alt, _, _ = zero_filled_rss(k, m, synth.CROP, None, schedule="pair", chunk_slices=chunk)
        assert torch.equal(alt, out), chunk
    nout, mean, std = zero_filled_rss(k, m, synth.CROP, "instance", schedule="pair")
    nref, rmean, rstd = zero_filled_rss(k, m, synth.CROP, "instance", schedule="sequential")
    torch.testing.assert_close(nout, nref, rtol=0, atol=2e-5)
    torch.testing.assert_close(mean, rmean, rtol=1e-5, atol=0)
    torch.testing.assert_close(std, rstd, rtol=1e-5, atol=0)
    k2 = k.clone()
    k2[..., torch.from_numpy(m == 0).cuda()] = complex(1e30, -1e30)
    alt, _, _ = zero_filled_rss(k2, m, synth.CROP, None, schedule="pair")
    assert torch.equal(alt, out)
    m2 = np.zeros(368, np.float32); m2[1::4] = 1.0; m2[180:190] = 0.5
    a, _, _ = zero_filled_rss(k, m2, (77, 200), None, schedule="pair")
    b, _, _ = zero_filled_rss(k, m2, (77, 200), None, schedule="sequential")
    assert O.rel_l2(a.cpu().numpy(), b.cpu().numpy()) <= 2e-6
    m3 = np.zeros(368, np.float32); m3[0::4] = 1.0; m3[[1, 17, 33]] = 1.0
    a, _, _ = zero_filled_rss(k, m3, synth.CROP, None, schedule="pair")
    b, _, _ = zero_filled_rss(k, m3, synth.CROP, None, schedule="sequential")
    assert torch.equal(a, b)


def test_other_widths_behind_the_640_column_pass():
    """640 x 372 knee files (31 x 12 plan of the 16-row row pass) and 640 x 320 / 400 (no specialised row kernel: pruned
    generic row pass), against the oracle's numpy chain and the full generic path; chunking never changes a slice."""
    for W, seed in ((372, 61), (320, 62), (400, 63)):
        k_np = synth.gaussian_kspace((3, 15, 640, W), seed)
        k = torch.from_numpy(k_np).cuda()
        m = synth.equispaced_mask(W, 4, 0.08)
        out, mean, std = zero_filled_rss(k, m, (320, 320), "instance")
        want, wmean, wstd = O.knee_chain_numpy(k_np[1], m, (320, 320), "instance")
        assert O.rel_l2(out[1].cpu().numpy(), want) <= TOL
        np.testing.assert_allclose([float(mean[1]), float(std[1])], [float(wmean), float(wstd)], rtol=1e-5)
        gen, _, _ = zero_filled_rss(k, m, (320, 320), "instance", force_generic=True)
        assert O.rel_l2(out.cpu().numpy(), gen.cpu().numpy()) <= 2e-6
        c1, _, _ = zero_filled_rss(k, m, (320, 320), "instance", chunk_slices=1)
        assert torch.equal(c1, out)


def test_wide_lines_fall_back_to_generic_kernels():
    """640 x 4096 / 640 x 3072: one line of the pruned generic row pass would need 292 / 245 KB of shared memory, so the
    call must take the generic kernels (which hold lines up to 8192) instead of failing -- and `supported` says GENERIC."""
    from mri_acl_imagesegmentation_adsp_b200.adapters import recon_cabi
    for W, seed in ((4096, 81), (3072, 82)):
        assert recon_cabi.library().supported(640, W) == recon_cabi.PATH_GENERIC
        k_np = synth.gaussian_kspace((1, 2, 640, W), seed)
        m = synth.equispaced_mask(W, 4, 0.08)
        out, _, _ = zero_filled_rss(torch.from_numpy(k_np).cuda(), m, (320, 320), None)
        want, _, _ = O.knee_chain_numpy(k_np[0], m, (320, 320), None)
        assert O.rel_l2(out[0].cpu().numpy(), want) <= TOL


def test_plan_cache_is_bounded():
    """a loader that draws a new random mask per volume: more distinct masks than the plan cache keeps (64); evicted plans
    are rebuilt on demand and results never change."""
    rng = np.random.default_rng(77)
    k = torch.from_numpy(synth.gaussian_kspace((1, 2, 640, 368), 71)).cuda()
    masks = [(rng.uniform(size=368) < 0.3).astype(np.float32) for _ in range(70)]
    first, _, _ = zero_filled_rss(k, masks[0], synth.CROP, None)
    ref0 = first.clone()
    for m in masks[1:]:
        zero_filled_rss(k, m, synth.CROP, None)
    again, _, _ = zero_filled_rss(k, masks[0], synth.CROP, None)       # its plan was evicted in between
    assert torch.equal(again, ref0)
    want = O.center_crop(np.sqrt((O.complex_abs(O.ifft2c(O.apply_mask(k[0].cpu().numpy(), masks[69]))) ** 2).sum(0)), synth.CROP)
    last, _, _ = zero_filled_rss(k, masks[69], synth.CROP, None)
    assert O.rel_l2(last[0].cpu().numpy(), want.astype(np.float32)) <= TOL


def test_fused_variants_against_oracle():
    rng = np.random.default_rng(5)
    k = synth.gaussian_kspace((2, 2, 3, 640, 368), 7)       # (S, A, C, H, W)
    m = np.zeros(368, np.float32)
    idx = np.sort(rng.choice(368, size=21, replace=False))
    m[idx] = rng.uniform(0.5, 1.5, size=21).astype(np.float32)
    out, mean, std = zero_filled_rss(k, m, (77, 200), "instance", flip_rows=True, average_axis=1, chunk_slices=1)
    gen, gmean, gstd = zero_filled_rss(k, m, (77, 200), "instance", flip_rows=True, average_axis=1, force_generic=True)
    for s in range(2):
        ims = []
        for a in range(2):
            mag = O.complex_abs(O.ifft2c(O.apply_mask(k[s, a], m)))
            ims.append(np.flipud(np.sqrt((mag ** 2).sum(0))))
        ref = np.ascontiguousarray(O.center_crop(np.mean(ims, axis=0), (77, 200)).astype(np.float32))
        nref, rmean, rstd = O.normalize_instance(ref)
        assert O.rel_l2(out[s], nref) <= TOL
        assert O.rel_l2(gen[s], nref) <= TOL
        np.testing.assert_allclose([mean[s], std[s]], [rmean, rstd], rtol=1e-5)
        np.testing.assert_allclose([gmean[s], gstd[s]], [rmean, rstd], rtol=1e-5)


def test_single_coil_live_path(golden):
    k = synth.gaussian_kspace((640, 368), 1)
    a = MRIKneePreprocessor.ifft2c_single(k)
    assert a.dtype == np.float32 and a.shape == (640, 368)
    assert O.rel_l2(a[::2], golden["single_coil/ifft2c_single_rows_even"]) <= TOL
    b = MRIKneePreprocessor.ifft2c_single(synth.gaussian_kspace((640, 372), 2))     # 372-wide file: generic kernels
    assert O.rel_l2(b[::5, ::3], golden["single_coil/ifft2c_single_372_sub"]) <= TOL
    recs = [{"kspace": synth.gaussian_kspace((640, 368), 50 + i), "meta": {"slice": i}} for i in range(3)]
    pack = MRIKneePreprocessor().recon_records(recs)
    assert pack["tensor"].shape == (3, 1, 640, 368) and pack["tensor"].dtype == torch.float32
    for i, r in enumerate(recs):
        assert O.rel_l2(pack["tensor"][i, 0].cpu().numpy(), O.ifft2c_single(r["kspace"])) <= TOL
    with pytest.raises(ValueError):
        MRIKneePreprocessor().recon_records([{"kspace": np.zeros((2, 8, 8), np.float32)}])


def test_analytic_and_parseval():
    k = np.zeros((1, 640, 368), np.complex64)
    k[0, 320, 184] = 1.0
    img, _, _ = zero_filled_rss(k, None, None, None)
    np.testing.assert_allclose(img, 1.0 / np.sqrt(640 * 368), rtol=2e-6)
    # batch 64 at full size, device-generated: Parseval per slice (sum RSS^2 == sum |k|^2)
    g = torch.Generator(device="cuda").manual_seed(1)
    kb = torch.view_as_complex(torch.randn((64, 15, 640, 368, 2), device="cuda", generator=g))
    full, _, _ = zero_filled_rss(kb, None, None, None)
    e_img = (full.double() ** 2).sum(dim=(1, 2))
    e_k = (torch.view_as_real(kb).double() ** 2).sum(dim=(1, 2, 3, 4))
    assert float(((e_img - e_k).abs() / e_k).max()) <= 1e-5
    # linearity in magnitude: scaling k-space by c scales the un-normalised image by |c|,
    # and leaves the instance-normalised image unchanged
    m = synth.knee_mask()
    a, _, _ = zero_filled_rss(kb[:4], m, synth.CROP, None)
    b, _, _ = zero_filled_rss(kb[:4] * 3.0, m, synth.CROP, None)
    assert float((b - 3.0 * a).norm() / (3.0 * a).norm()) <= 2e-6
    na, mean, std = zero_filled_rss(kb[:4], m, synth.CROP, "instance")
    nb, _, _ = zero_filled_rss(kb[:4] * 3.0, m, synth.CROP, "instance")
    assert float((na - nb).norm() / na.norm()) <= TOL
    # the normalised output has zero mean and unit (unbiased) std
    assert float(na.mean(dim=(1, 2)).abs().max()) <= 1e-5
    assert float((na.std(dim=(1, 2)) - 1).abs().max()) <= 1e-5
    torch.testing.assert_close(mean, a.mean(dim=(1, 2)), rtol=1e-5, atol=0)
    torch.testing.assert_close(std, a.std(dim=(1, 2)), rtol=1e-5, atol=0)


def test_layouts_and_types():
    k = synth.gaussian_kspace((2, 4, 640, 368), 60)
    m = synth.knee_mask()
    ref, _, _ = zero_filled_rss(k, m)
    t_cpu = torch.from_numpy(k)
    o_cpu, _, _ = zero_filled_rss(t_cpu, torch.from_numpy(m))
    assert isinstance(o_cpu, torch.Tensor) and o_cpu.device.type == "cpu"
    np.testing.assert_array_equal(o_cpu.numpy(), ref)
    o_ri, _, _ = zero_filled_rss(torch.view_as_real(t_cpu).cuda(), m)        # fastMRI real view (...,2)
    assert o_ri.is_cuda
    np.testing.assert_array_equal(o_ri.cpu().numpy(), ref)
    nc = torch.from_numpy(k).cuda().permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2)   # non-contiguous view
    o_nc, _, _ = zero_filled_rss(nc, m)
    np.testing.assert_array_equal(o_nc.cpu().numpy(), ref)
    x = recon_to_unet_input(torch.from_numpy(k).cuda(), m)
    assert x.shape == (2, 1, 320, 320) and x.dtype == torch.float32 and x.is_contiguous()
    empty, _, _ = zero_filled_rss(torch.zeros((0, 4, 640, 368), dtype=torch.complex64, device="cuda"), m)
    assert empty.shape == (0, 320, 320)


def test_errors():
    k = synth.gaussian_kspace((2, 16, 12), 1)
    with pytest.raises(ValueError):
        zero_filled_rss(k, None, (17, 12))
    with pytest.raises(ValueError):
        zero_filled_rss(k, np.ones(11, np.float32), (8, 8))
    with pytest.raises(ValueError):
        zero_filled_rss(k[0], None, (8, 8))
    with pytest.raises(ValueError):
        zero_filled_rss(np.zeros((2, 16, 12), np.float32), None, (8, 8))
    with pytest.raises(ValueError):
        zero_filled_rss(k, None, (8, 8), normalize="zscore")
    with pytest.raises(ValueError):
        K.ifft2c(np.zeros((1, 5000, 4), np.complex64))


# ---------------------------------------------------------------------------------------------
# prostate T2 chain, configs[2]
# ---------------------------------------------------------------------------------------------
def test_prostate_small(golden):
    k = golden["prostate_small/kspace"]
    assert t2.padding_lr(32, 20) == (5, 6) and t2.padding_lr(640, 450) == (94, 95)
    for av in range(2):
        cc = t2.create_coil_combined_im(O.zero_pad_pe(k[av], 5, 6))
        assert cc.dtype == np.float64
        assert O.rel_l2(cc, golden["prostate_small/coil_combined"][av]) <= TOL
    fin = t2.t2_average_combine(k, (5, 6), (16, 16))
    assert fin.dtype == np.float64 and fin.shape == (2, 16, 16)
    assert O.rel_l2(fin, golden["prostate_small/final_16x16"]) <= TOL


def test_prostate_config2_one_slice(golden, manifest):
    case = manifest["cases"]["prostate_one_slice"]
    k = synth.gaussian_kspace(tuple(case["shape"]), case["seed"])
    fin = t2.t2_average_combine(k, synth.PROSTATE_PAD, synth.CROP, mask=synth.prostate_mask())
    assert fin.shape == (1, 320, 320)
    assert O.rel_l2(fin, golden["prostate/one_slice_final"]) <= TOL


def test_prostate_config2_full_volume(golden):
    """configs[2] at its own shape (3, 30, 16, 640, 451): slice stride x average stride x 8x mask x pad (94, 95) x flipud x
    chunking all at once through the fused 640-wide plan, against the frozen reference outputs (slices 0 and 29), the
    oracle (two more slices) and itself under chunk_slices 1 / 7 / 30 (bit-equal)."""
    A, S, C, RO, PE = synth.PROSTATE_SHAPE
    k = torch.empty((A, S, C, RO, PE), dtype=torch.complex64, device="cuda")
    for a in range(A):
        for s in range(S):
            k[a, s] = torch.from_numpy(synth.prostate_volume_block(a, s))
    pm = synth.prostate_mask()
    fin = t2.t2_average_combine(k, synth.PROSTATE_PAD, synth.CROP, mask=pm)
    assert fin.shape == (S, 320, 320) and fin.dtype == torch.float64 and fin.is_cuda
    for s in (0, 29):
        assert O.rel_l2(fin[s, ::2, ::2].cpu().numpy(), golden[f"prostate/volume_slice{s}_sub2"]) <= TOL, s
    for s in (7, 18):
        want = O.prostate_chain(k[:, s:s + 1].cpu().numpy(), pm, synth.PROSTATE_PAD, synth.CROP)
        assert O.rel_l2(fin[s].cpu().numpy(), want[0]) <= TOL, s
    ref, _, _ = zero_filled_rss(k, pm, synth.CROP, None, flip_rows=True, average_axis=0, pad=synth.PROSTATE_PAD, chunk_slices=30)
    assert torch.equal(ref.double(), fin)
    for chunk in (1, 7):
        alt, _, _ = zero_filled_rss(k, pm, synth.CROP, None, flip_rows=True, average_axis=0, pad=synth.PROSTATE_PAD,
                                    chunk_slices=chunk)
        assert torch.equal(alt, ref), chunk


def test_batch64_config1_against_oracle():
    """the benchmarked shape itself: slices 0 / 31 / 63 of a batch-64 configs[1] call against the oracle's numpy chain."""
    g = torch.Generator(device="cuda").manual_seed(1234)
    kb = torch.view_as_complex(torch.randn((64,) + synth.KNEE_SHAPE + (2,), device="cuda", generator=g))
    m = synth.knee_mask()
    img, mean, std = zero_filled_rss(kb, m, synth.CROP, "instance")
    assert img.shape == (64, 320, 320)
    for s in (0, 31, 63):
        want, wmean, wstd = O.knee_chain_numpy(kb[s].cpu().numpy(), m, synth.CROP, "instance")
        assert O.rel_l2(img[s].cpu().numpy(), want) <= TOL, s
        np.testing.assert_allclose([float(mean[s]), float(std[s])], [float(wmean), float(wstd)], rtol=1e-5)
    one, _, _ = zero_filled_rss(kb[31], m, synth.CROP, "instance")
    assert torch.equal(one, img[31])


@pytest.mark.parametrize("tag,spec", [("8x", (368, 8, 0.04, 0)), ("4x_off1", (368, 4, 0.08, 1)), ("4x_off3", (368, 4, 0.08, 3))])
def test_knee_other_masks_default_schedule(golden, manifest, tag, spec):
    """the 8x knee mask (60 of 368 columns) and 4x masks with a non-zero offset through the DEFAULT schedule, against
    the frozen reference outputs; the mask index lists themselves are frozen in the manifest."""
    m = synth.equispaced_mask(*spec)
    assert np.flatnonzero(m).tolist() == manifest["masks"][synth.mask_name(*spec)]
    k = synth.gaussian_kspace(synth.KNEE_SHAPE, 0)
    raw, _, _ = zero_filled_rss(k, m, synth.CROP, None)
    assert O.rel_l2(raw[::2, ::2], golden[f"knee_gauss/numpy_chain_{tag}_sub2"]) <= TOL
    img, mean, std = zero_filled_rss(k, m, synth.CROP, "instance")
    np.testing.assert_allclose([mean, std], golden[f"knee_gauss/fastmri_mean_std_{tag}"], rtol=1e-5)
    want, _, _ = O.knee_chain_numpy(k, m, synth.CROP, "instance")
    assert O.rel_l2(img, want) <= TOL
    k2 = k.copy()
    k2[..., m == 0] = complex(1e30, -1e30)           # masked columns are never read
    raw2, _, _ = zero_filled_rss(k2, m, synth.CROP, None)
    np.testing.assert_array_equal(raw2, raw)
    kb = torch.from_numpy(np.stack([k, k2, k])).cuda()   # a batch through the same plan, chunked
    rb, _, _ = zero_filled_rss(kb, m, synth.CROP, None, chunk_slices=2)
    for i in range(3):
        np.testing.assert_array_equal(rb[i].cpu().numpy(), raw)


def test_reference_signature_twins():
    """the remaining reference signatures: ifftnd with its default axes=[-1] / any axes / None,
    flip_im, center_crop_im, numpy rss, complex_center_crop, center_crop_to_smallest; dtype follows the input
    (complex128 -> complex128 / float64) like the numpy functions they replace."""
    x = synth.gaussian_kspace((3, 30, 23), 81)
    for axes in ((-1,), [0], [1], [0, 2], [1, 2], None):
        got = t2.ifftnd(x, axes) if axes != (-1,) else t2.ifftnd(x)
        assert got.dtype == np.complex64 and got.shape == x.shape
        assert O.rel_l2(got, O.ifftnd(x.copy(), list(axes) if axes is not None else None)) <= TOL, axes
    x128 = x.astype(np.complex128)
    g128 = t2.ifftnd(x128, [1, 2])
    assert g128.dtype == np.complex128 and O.rel_l2(g128, O.ifftnd(x128.copy(), [1, 2])) <= TOL
    assert K.ifft2c(x128).dtype == np.complex128 and K.fft2c(x128).dtype == np.complex128
    assert K.complex_abs(x128).dtype == np.float64 and K.complex_abs(x).dtype == np.float32
    tt = torch.from_numpy(x128).cuda()
    assert K.ifft2c(tt).dtype == torch.complex128 and K.ifft2c(tt).is_cuda
    gt = t2.ifftnd(torch.from_numpy(x).cuda(), [0])
    assert gt.is_cuda and O.rel_l2(gt.cpu().numpy(), O.ifftnd(x.copy(), [0])) <= TOL
    with pytest.raises(ValueError):
        t2.ifftnd(x, [1, 1])
    vol = np.abs(x).astype(np.float64)
    for ax in (0, 1):
        np.testing.assert_array_equal(t2.flip_im(vol[:, :3].copy(), ax), O.flip_im(vol[:, :3].copy(), ax))
    tv = torch.from_numpy(vol[:, :3].copy())
    np.testing.assert_array_equal(t2.flip_im(tv, 0).numpy(), O.flip_im(vol[:, :3].copy(), 0))
    np.testing.assert_array_equal(t2.center_crop_im(vol, [7, 12]), O.center_crop_im(vol, [7, 12]))
    for ax in (-1, 0, 1):
        r = t2.rss(x, ax)
        assert r.dtype == np.float32 and O.rel_l2(r, O.rss_np(x, ax)) <= 1e-6
    assert t2.rss(x128, 0).dtype == np.float64
    assert O.rel_l2(t2.rss(np.abs(x), 0), O.rss_np(np.abs(x), 0)) <= 1e-6
    t = transforms.to_tensor(x)
    np.testing.assert_array_equal(transforms.complex_center_crop(t, (12, 9)).numpy(), O.complex_center_crop_ri(t.numpy(), (12, 9)))
    with pytest.raises(ValueError):
        transforms.complex_center_crop(t, (31, 9))
    a, b = torch.from_numpy(vol[:, :8, :]), torch.from_numpy(vol[:, :, :5])
    ra, rb = transforms.center_crop_to_smallest(a, b)
    oa, ob = O.center_crop_to_smallest(a.numpy(), b.numpy())
    np.testing.assert_array_equal(ra.numpy(), oa)
    np.testing.assert_array_equal(rb.numpy(), ob)


_EXP_LIB = os.path.join(os.path.dirname(cabi.DEFAULT_LIBRARY), "libmriacl_recon_exp.so")


@pytest.mark.skipif(not os.path.exists(_EXP_LIB), reason="experimental library not built (make -C csrc experimental)")
@pytest.mark.parametrize("env", [
    {"MRIACL_CP_TMA": "2", "MRIACL_CT_SLOTS": "4"},                         # stand-alone TMA column pass, two teams
    {"MRIACL_CP_TMA": "1"},                                                 # one team per CTA, two CTAs per SM
    {"MRIACL_SCHEDULE": "coresident", "MRIACL_KC_TMA": "1"},                # TMA column teams + row team in one CTA
    {"MRIACL_SCHEDULE": "coresident", "MRIACL_KC_TMA": "1", "MRIACL_KC_RING": "12"},
    {"MRIACL_SCHEDULE": "coresident", "MRIACL_KC_SPLIT_SM": "64"},          # column SMs / row SMs, ring of 16 T slots
], ids=["tma2", "tma1", "cores_tma", "cores_tma_ring", "sm_split"])
def test_tma_experimental_schedules(env):
    """the TMA band gather (colpass640_tma.cuh) under the three schedules built on it gives the oracle's images: the
    schedules read their knobs once per process, so each runs in its own interpreter (tests/_tma_schedule_check.py)."""
    e = dict(os.environ, MRIACL_RECON_LIBRARY=_EXP_LIB, **env)
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "_tma_schedule_check.py")],
                       env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().startswith("OK"), r.stdout[-2000:] + r.stderr[-4000:]


def test_640_wide_balanced_first_pass_mixed_kinds():
    """the balanced first pass of rowpass640_kernel<true> (at most 60 of the 80 butterfly positions hold samples) with both
    kinds of unit in one plan: positions with all eight inputs and positions with three weighted ones; several slices and
    coils so that every CTA takes more than one item."""
    k = synth.gaussian_kspace((3, 4, 640, 640), 43)
    m = np.zeros(640, np.float32)
    for pos in range(0, 10):
        for n1 in range(8):
            m[(80 * n1 + pos + 320) % 640] = 1.0
    for pos in range(20, 45):
        for n1 in (1, 4, 6):
            m[(80 * n1 + pos + 320) % 640] = 0.5 + 0.01 * pos
    img, mean, std = zero_filled_rss(torch.from_numpy(k).cuda(), m, synth.CROP, "instance")
    ref, rmean, rstd = O.knee_chain_numpy(k, m, synth.CROP, "instance")
    for s in range(3):
        assert O.rel_l2(img[s].cpu().numpy(), ref[s]) <= TOL
    np.testing.assert_allclose(mean.cpu().numpy(), rmean, rtol=1e-5)
    gen, _, _ = zero_filled_rss(torch.from_numpy(k).cuda(), m, synth.CROP, "instance", force_generic=True)
    torch.testing.assert_close(gen, img, rtol=0, atol=2e-5)
