"""The kernel sources and host orchestration of libmriacl_recon, compiled for the CPU with the
thread-per-CUDA-thread emulator in tests/emu/ and checked against the oracle.

This is how index maps, plans, barriers and the C-ABI argument handling are validated in the
GPU-less build container.  It is test infrastructure: the emulation library is built into
tests/emu/_build/ and is never loaded by the product package.  The `-m gpu` tests run the same
checks (and more) on the real sm_100a build.
"""
import os
import subprocess

import numpy as np
import pytest

from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.adapters import recon_cabi as cabi
from oracle import recon_oracle as O

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
TOL = 1e-5   # BASELINE.json: rel-L2 <= 1e-5 in fp32


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", EMU_DIR], check=True)
    return cabi.ReconLibrary(os.path.join(EMU_DIR, "_build", "libmriacl_emu.so"))


def P(a):
    return a.ctypes.data


def fft2c(lib, x, inverse):
    x = np.ascontiguousarray(x)
    out = np.empty_like(x)
    b = int(np.prod(x.shape[:-2])) if x.ndim > 2 else 1
    lib.fft2c(P(x), P(out), b, x.shape[-2], x.shape[-1], inverse)
    return out


def recon(lib, k, mask, crop, flags=0, pad=(0, 0), eps=0.0, chunk=None):
    """k: (S, A, C, H, W) complex64"""
    S, A, C, H, W = k.shape
    Wp = W + pad[0] + pad[1]
    out = np.zeros((S,) + tuple(crop), np.float32)
    ms = np.zeros((S, 2), np.float32)
    nbytes = lib.recon_rss_workspace_bytes(chunk or S, A, C, H, W, pad[0], Wp, crop[0], crop[1], mask, flags)
    ws = np.zeros(nbytes, np.uint8)
    lib.recon_rss(P(k), A * C * H * W, C * H * W, mask, P(out), P(ms), S, A, C, H, W, pad[0], Wp, crop[0], crop[1],
                  flags, eps, P(ws), nbytes)
    return out, ms


@pytest.mark.parametrize("shape", [(2, 16, 12), (3, 15, 11), (1, 30, 23), (1, 37, 5), (1, 64, 40)])
def test_generic_fft2c(emu, shape):
    x = synth.gaussian_kspace(shape, 3)
    assert O.rel_l2(fft2c(emu, x, True), O.ifft2c(x)) <= TOL
    assert O.rel_l2(fft2c(emu, x, False), O.fft2c(x)) <= TOL


def test_generic_chain_small(emu, golden):
    k, m = golden["small_even/kspace"], golden["small_even/mask"]
    out, ms = recon(emu, k[None, None], m, (16, 16), cabi.NORM_INSTANCE)
    assert O.rel_l2(out[0], golden["small_even/fastmri_chain_16x16"]) <= TOL
    np.testing.assert_allclose(ms[0], golden["small_even/fastmri_mean_std"], rtol=1e-5)
    raw, _ = recon(emu, k[None, None], m, (16, 16), 0)
    assert O.rel_l2(raw[0], golden["small_even/numpy_chain_16x16"]) <= TOL


def test_generic_prostate_small(emu, golden):
    k = golden["prostate_small/kspace"]                      # (A, S, C, RO, PE)
    kk = np.ascontiguousarray(k.transpose(1, 0, 2, 3, 4))     # (S, A, C, RO, PE)
    out, _ = recon(emu, kk, None, (16, 16), cabi.FLIP_ROWS, pad=(5, 6))
    assert O.rel_l2(out, golden["prostate_small/final_16x16"]) <= TOL


def test_fused_knee_slice(emu, golden):
    """configs[0]: the fused column pass + row pass + normalise on the 15-coil 640x368 slice."""
    assert emu.supported(640, 368) == cabi.PATH_FUSED and emu.supported(640, 320) == cabi.PATH_GENERIC
    k = synth.gaussian_kspace((1, 1) + synth.KNEE_SHAPE, 0)
    m = synth.knee_mask()
    out, ms = recon(emu, k, m, synth.CROP, cabi.NORM_INSTANCE)
    assert O.rel_l2(out[0], golden["knee_gauss/fastmri_chain"]) <= TOL
    np.testing.assert_allclose(ms[0], golden["knee_gauss/fastmri_mean_std"], rtol=1e-5)
    raw, _ = recon(emu, k, m, synth.CROP, 0)
    assert O.rel_l2(raw[0], golden["knee_gauss/numpy_chain"]) <= TOL
    # crop indexing is exact: the cropped launch equals the window of the uncropped one, bit for bit
    full, _ = recon(emu, k, m, (640, 368), 0)
    np.testing.assert_array_equal(raw[0], full[0, 160:480, 24:344])
    # masked-out columns are never read: garbage there changes nothing, bit for bit
    k2 = k.copy()
    k2[..., m == 0] = np.complex64(1e30 + 1e30j)
    raw2, _ = recon(emu, k2, m, synth.CROP, 0)
    np.testing.assert_array_equal(raw, raw2)


def test_fused_variants(emu):
    """ragged column groups, flip, averages, odd crop, weighted mask, two slices in two chunks."""
    rng = np.random.default_rng(5)
    k = synth.gaussian_kspace((2, 2, 2, 640, 368), 7)
    m = np.zeros(368, np.float32)
    idx = np.sort(rng.choice(368, size=21, replace=False))
    m[idx] = rng.uniform(0.5, 1.5, size=21).astype(np.float32)
    out, ms = recon(emu, k, m, (77, 200), cabi.FLIP_ROWS | cabi.NORM_INSTANCE, chunk=1)
    for s in range(2):
        ims = []
        for a in range(2):
            img = O.complex_abs(O.ifft2c(O.apply_mask(k[s, a], m)))
            ims.append(np.flipud(np.sqrt((img ** 2).sum(0))))
        ref = O.center_crop(np.mean(ims, axis=0), (77, 200)).astype(np.float32)
        nref, mean, std = O.normalize_instance(np.ascontiguousarray(ref))
        assert O.rel_l2(out[s], nref) <= TOL
        np.testing.assert_allclose(ms[s], [mean, std], rtol=1e-5)
    gen, _ = recon(emu, k, m, (77, 200), cabi.FLIP_ROWS | cabi.NORM_INSTANCE | cabi.FORCE_GENERIC)
    assert O.rel_l2(out, gen) <= TOL


def test_fused_schedules_agree(emu):
    """the persistent fused kernel (dynamic claiming of column / row items) and the overlapped schedule's kernels
    produce the same bits as the sequential schedule (the emulator runs the launches one after another)."""
    k = synth.gaussian_kspace((2, 1, 3, 640, 368), 17)
    m = synth.knee_mask()
    ref, _ = recon(emu, k, m, (320, 320), cabi.SEQUENTIAL)
    for flag in (cabi.SCHED_FUSED, cabi.SCHED_OVERLAP):
        out, _ = recon(emu, k, m, (320, 320), flag)
        np.testing.assert_array_equal(out, ref)
        out1, _ = recon(emu, k, m, (320, 320), flag, chunk=1)
        np.testing.assert_array_equal(out1, ref)


def test_coresident_kernel(emu):
    """column team + row team in one CTA, per-slice counters, normalisation by the team that finishes a slice's
    last tile: same bits as the back-to-back kernels for the images, same statistics."""
    k = synth.gaussian_kspace((2, 2, 2, 640, 368), 29)
    m = synth.knee_mask()
    ref, rms = recon(emu, k, m, (320, 320), cabi.SEQUENTIAL | cabi.NORM_INSTANCE)
    out, ms = recon(emu, k, m, (320, 320), cabi.SCHED_CORESIDENT | cabi.NORM_INSTANCE)
    np.testing.assert_allclose(out, ref, rtol=0, atol=2e-6)
    np.testing.assert_allclose(ms, rms, rtol=1e-6)
    raw_ref, _ = recon(emu, k, m, (77, 200), cabi.SEQUENTIAL | cabi.FLIP_ROWS)
    raw, _ = recon(emu, k, m, (77, 200), cabi.SCHED_CORESIDENT | cabi.FLIP_ROWS, chunk=1)
    np.testing.assert_array_equal(raw, raw_ref)
    # single average + a mask of the pair family: the row team is the pair row pass (rowpair.cuh)
    k1 = synth.gaussian_kspace((2, 1, 3, 640, 368), 37)
    ref1, rms1 = recon(emu, k1, m, (320, 320), cabi.SEQUENTIAL | cabi.NORM_INSTANCE)
    out1, ms1 = recon(emu, k1, m, (320, 320), cabi.SCHED_CORESIDENT | cabi.NORM_INSTANCE)
    assert O.rel_l2(out1, ref1) <= TOL
    np.testing.assert_allclose(ms1, rms1, rtol=1e-5)


def test_pipelined_schedule_and_fused_normalisation(emu):
    """chunk-pipelined schedule (two T buffers, side stream) and the normalisation fused into the row pass (last tile
    of a slice normalises it) against the separate normalise launch."""
    import os
    k = synth.gaussian_kspace((3, 1, 2, 640, 368), 31)
    m = synth.knee_mask()
    os.environ["MRIACL_FUSE_NORM"] = "1"       # opt-in knob, read per call
    ref_raw, _ = recon(emu, k, m, (320, 320), cabi.SEQUENTIAL)
    out, ms = recon(emu, k, m, (320, 320), cabi.SEQUENTIAL | cabi.NORM_INSTANCE)
    for s in range(3):
        nref, mean, std = O.normalize_instance(np.ascontiguousarray(ref_raw[s]))
        assert O.rel_l2(out[s], nref) <= TOL
        np.testing.assert_allclose(ms[s], [mean, std], rtol=1e-5)
    pout, pms = recon(emu, k, m, (320, 320), cabi.SCHED_PIPELINED | cabi.NORM_INSTANCE, chunk=2)
    np.testing.assert_array_equal(pout, out)
    np.testing.assert_array_equal(pms, ms)
    os.environ.pop("MRIACL_FUSE_NORM")
    sep, sms_ = recon(emu, k, m, (320, 320), cabi.SEQUENTIAL | cabi.NORM_INSTANCE)
    np.testing.assert_allclose(sep, out, rtol=0, atol=2e-6)
    # co-resident schedule with a one-slot ring of T: the column teams wait for the row teams before reusing the slot
    os.environ["MRIACL_KC_RING"] = "1"
    rout, rms_ = recon(emu, k, m, (320, 320), cabi.SCHED_CORESIDENT | cabi.NORM_INSTANCE)
    os.environ.pop("MRIACL_KC_RING")
    assert O.rel_l2(rout, out) <= TOL


def test_fused_640_wide_plan(emu):
    """prostate-shape plan (PE 451 zero-padded to 640, 8x mask + ACS, flipud, mean over averages after the RSS, crop):
    column pass + 640-wide row pass (expanding first pass driven by the per-position plan) against the oracle; also
    fully sampled (regular butterflies everywhere) and an irregular mask (partial butterflies), normalisation included."""
    rng = np.random.default_rng(11)
    k = synth.gaussian_kspace((1, 2, 2, 640, 451), 41)           # (S, A, C, RO, PE)

    def oracle(mask, crop):
        ims = []
        for a in range(2):
            kk = np.pad(O.apply_mask(k[0, a], mask), ((0, 0), (0, 0), (94, 95)))
            ims.append(np.flipud(np.sqrt((O.complex_abs(O.ifft2c(kk)) ** 2).sum(0))))
        return O.center_crop(np.mean(ims, axis=0), crop).astype(np.float32)

    for m in (synth.prostate_mask(), None, (rng.uniform(size=451) < 0.3).astype(np.float32)):
        out, ms = recon(emu, k, m, (320, 320), cabi.FLIP_ROWS | cabi.NORM_INSTANCE, pad=(94, 95))
        nref, mean, std = O.normalize_instance(np.ascontiguousarray(oracle(m, (320, 320))))
        assert O.rel_l2(out[0], nref) <= TOL
        np.testing.assert_allclose(ms[0], [mean, std], rtol=1e-5)
    raw, _ = recon(emu, k, synth.prostate_mask(), (77, 200), cabi.FLIP_ROWS, pad=(94, 95))
    assert O.rel_l2(raw[0], oracle(synth.prostate_mask(), (77, 200))) <= TOL


def test_640_wide_balanced_first_pass_mixed_kinds(emu):
    """640-wide plan without padding whose mask makes the balanced first pass (at most 60 of the 80 butterfly positions hold
    samples) meet both kinds of unit: positions with all eight inputs (regular radix-8) and positions with a few (direct sum),
    weighted columns included."""
    k = synth.gaussian_kspace((1, 1, 2, 640, 640), 43)
    m = np.zeros(640, np.float32)
    for pos in range(0, 10):                       # logical n = 80 n1 + pos for every n1: full positions
        for n1 in range(8):
            m[(80 * n1 + pos + 320) % 640] = 1.0
    for pos in range(20, 45):                      # three inputs each, weighted
        for n1 in (1, 4, 6):
            m[(80 * n1 + pos + 320) % 640] = 0.5 + 0.01 * pos
    out, _ = recon(emu, k, m, (48, 320), 0)
    kk = O.apply_mask(k[0, 0], m)
    want = O.center_crop(np.sqrt((O.complex_abs(O.ifft2c(kk)) ** 2).sum(0)), (48, 320)).astype(np.float32)
    assert O.rel_l2(out[0], want) <= TOL


@pytest.mark.parametrize("W", [372, 400, 320])
def test_other_knee_widths(emu, W):
    """H = 640 with the other knee widths: 372 = 31 x 12 and 400 = 25 x 16 run the 16-row row pass with a 31- / 25-point
    first stage and a 12- / 16-point second stage; 320 has no specialised row kernel and takes the pruned generic row
    pass (Stockham on the kept rows only, RSS / averages / crop fused).  Against the oracle, normalisation included;
    then fully sampled, flipped, with a ragged last tile and the full width."""
    assert emu.supported(640, W) == (cabi.PATH_GENERIC if W == 320 else cabi.PATH_FUSED)
    k = synth.gaussian_kspace((1, 2, 2, 640, W), 43)
    m = synth.equispaced_mask(W, 4, 0.08)
    crop = (320, min(320, W))
    out, ms = recon(emu, k, m, crop, cabi.NORM_INSTANCE)
    ims = [np.sqrt((O.complex_abs(O.ifft2c(O.apply_mask(k[0, a], m))) ** 2).sum(0)) for a in range(2)]
    nref, mean, std = O.normalize_instance(np.ascontiguousarray(O.center_crop(np.mean(ims, axis=0), crop).astype(np.float32)))
    assert O.rel_l2(out[0], nref) <= TOL
    np.testing.assert_allclose(ms[0], [mean, std], rtol=1e-5)
    raw, _ = recon(emu, k, None, (75, W), cabi.FLIP_ROWS)
    ims = [np.flipud(np.sqrt((O.complex_abs(O.ifft2c(k[0, a])) ** 2).sum(0))) for a in range(2)]
    ref = O.center_crop(np.mean(ims, axis=0), (75, W)).astype(np.float32)
    assert O.rel_l2(raw[0], ref) <= TOL


def test_pair_row_pass(emu, golden):
    """the pair row pass (rowpair.cuh: per-output-pair transform, stager warp, rotated residue-major tile) against
    the oracle and the cooperative row pass, incl. an odd crop, a shifted mask offset and a mask outside its family."""
    k = synth.gaussian_kspace((2, 1, 3, 640, 368), 23)
    m = synth.knee_mask()
    ref, rms = recon(emu, k, m, (320, 320), cabi.SEQUENTIAL | cabi.NORM_INSTANCE)
    out, ms = recon(emu, k, m, (320, 320), cabi.SCHED_PAIR | cabi.NORM_INSTANCE)
    assert O.rel_l2(out, ref) <= TOL
    np.testing.assert_allclose(ms, rms, rtol=1e-5)
    img = O.center_crop(O.rss(O.complex_abs(O.ifft2c(O.apply_mask(k[1, 0], m)))), (320, 320)).astype(np.float32)
    raw, _ = recon(emu, k, m, (320, 320), cabi.SCHED_PAIR, chunk=1)
    assert O.rel_l2(raw[1], img) <= TOL
    # equispaced offset 1 (dense residues need the index rotation), odd crop, weighted columns
    m2 = np.zeros(368, np.float32); m2[1::4] = 1.0; m2[180:190] = 0.5
    ref2, _ = recon(emu, k, m2, (77, 200), cabi.SEQUENTIAL)
    out2, _ = recon(emu, k, m2, (77, 200), cabi.SCHED_PAIR)
    assert O.rel_l2(out2, ref2) <= TOL
    # a mask outside the family (dense residues with gaps are fine; 3 extra columns in one residue are not): falls back
    m3 = np.zeros(368, np.float32); m3[0::4] = 1.0; m3[[1, 17, 33]] = 1.0
    ref3, _ = recon(emu, k, m3, (320, 320), cabi.SEQUENTIAL)
    out3, _ = recon(emu, k, m3, (320, 320), cabi.SCHED_PAIR)
    np.testing.assert_array_equal(out3, ref3)
    # thinned dense residues (zero slots)
    m4 = m.copy(); m4[[0, 8, 100, 364]] = 0.0
    ref4, _ = recon(emu, k, m4, (320, 320), cabi.SEQUENTIAL)
    out4, _ = recon(emu, k, m4, (320, 320), cabi.SCHED_PAIR)
    assert O.rel_l2(out4, ref4) <= TOL


def test_fused_single_coil_full(emu, golden):
    k = synth.gaussian_kspace((640, 368), 1)
    out = np.zeros((1, 640, 368), np.float32)
    nbytes = emu.ifft2c_abs_workspace_bytes(1, 640, 368)
    ws = np.zeros(nbytes, np.uint8)
    emu.ifft2c_abs(P(k), P(out), 1, 640, 368, P(ws), nbytes)
    assert O.rel_l2(out[0, ::2], golden["single_coil/ifft2c_single_rows_even"]) <= TOL


def test_elementwise_ops(emu, golden):
    k = golden["small_odd/kspace"]
    out = np.zeros(k.shape, np.float32)
    emu.complex_abs(P(k), P(out), k.size, False)
    assert O.rel_l2(out, O.complex_abs(k)) <= 1e-6
    img = np.ascontiguousarray(O.ifft2c(k))
    r = np.zeros((2, 30, 23), np.float32)
    emu.rss(P(img), P(r), 2, 3, 30 * 23, True)
    assert O.rel_l2(r, golden["small_odd/rss_complex"]) <= TOL
    a = np.ascontiguousarray(np.abs(k))
    emu.rss(P(a), P(r), 2, 3, 30 * 23, False)
    assert O.rel_l2(r, golden["small_odd/rss_real"]) <= 1e-6
    x = golden["small_even/ifft2c_single_coil0"]
    for oh, ow in [(40, 12), (8, 8), (32, 24), (5, 50)]:
        o = np.zeros((oh, ow), np.float32)
        emu.center_crop_or_pad(P(x), P(o), 1, 32, 24, oh, ow, 4)
        np.testing.assert_array_equal(o, O.center_crop_or_pad(x, oh, ow))
    kc = np.ascontiguousarray(k[0])
    o = np.zeros((3, 12, 40), np.complex64)
    emu.center_crop_or_pad(P(kc), P(o), 3, 30, 23, 12, 40, 8)
    np.testing.assert_array_equal(o, O.center_crop_or_pad(kc, 12, 40))
    y = (np.abs(synth.gaussian_kspace((2, 320, 320), 3)) + 2).astype(np.float32)
    o = np.zeros_like(y)
    ms = np.zeros((2, 2), np.float32)
    emu.normalize_instance(P(y), P(o), P(ms), 2, 320 * 320, 1e-11)
    for i in range(2):
        ref, mean, std = O.normalize_instance(y[i], 1e-11)
        assert O.rel_l2(o[i], ref) <= 2e-6
        np.testing.assert_allclose(ms[i], [mean, std], rtol=2e-6)


def test_error_convention(emu):
    k = synth.gaussian_kspace((1, 1, 2, 16, 12), 1)
    with pytest.raises(ValueError):          # crop larger than the image -> "Invalid shapes."
        recon(emu, k, None, (17, 12))
    with pytest.raises(ValueError):
        emu.fft2c(P(k), P(k), 1, 0, 12, True)
    with pytest.raises(ValueError):
        emu.fft2c(P(k), P(k), 1, 5000, 12, True)
    out = np.zeros((1, 8, 8), np.float32)
    ws = np.zeros(64, np.uint8)
    with pytest.raises(RuntimeError):        # workspace too small
        emu.recon_rss(P(k), 2 * 16 * 12, 0, None, P(out), 0, 1, 1, 2, 16, 12, 0, 12, 8, 8, 0, 0.0, P(ws), 64)
    # empty batch is a no-op
    emu.recon_rss(P(k), 2 * 16 * 12, 0, None, P(out), 0, 0, 1, 2, 16, 12, 0, 12, 8, 8, 0, 0.0, P(ws), 64)
    # all-zero mask: zeros before normalisation (NaN after, as in the reference's 0/0)
    km = synth.gaussian_kspace((1, 1, 2, 640, 368), 2)
    z, _ = recon(emu, km, np.zeros(368, np.float32), (64, 64), 0)
    assert not z.any()
