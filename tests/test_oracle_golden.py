"""oracle/recon_oracle.py against the frozen outputs of the reference's own functions
(tests/golden/, made by oracle/make_golden.py).  CPU only.

Tolerances: the oracle issues the same numpy calls as the reference, so results are
normally bit-identical; pocketfft may pick a different SIMD path on another CPU, hence
rel-L2 <= 1e-6 (10x below the 1e-5 parity bar) instead of equality.  Index work is exact.
"""
import numpy as np
import pytest

from mri_acl_imagesegmentation_adsp_b200 import synth
from oracle import recon_oracle as O

TOL = 1e-6


def test_small_even_transforms(golden):
    k = golden["small_even/kspace"]
    assert O.rel_l2(O.ifft2c(k), golden["small_even/ifft2c"]) <= TOL
    assert O.rel_l2(O.fft2c(k), golden["small_even/fft2c"]) <= TOL
    assert O.ifft2c(k).dtype == np.complex64
    assert O.rel_l2(O.complex_abs(k), golden["small_even/complex_abs"]) <= TOL
    s = O.ifft2c_single(k[0])
    assert s.dtype == np.float32
    assert O.rel_l2(s, golden["small_even/ifft2c_single_coil0"]) <= TOL
    np.testing.assert_array_equal(O.center_crop_or_pad(golden["small_even/ifft2c_single_coil0"], 40, 12),
                                  golden["small_even/crop_or_pad_40x12"])


def test_small_even_chains(golden):
    k, m = golden["small_even/kspace"], golden["small_even/mask"]
    img, _, _ = O.knee_chain_numpy(k, m, (16, 16), normalize_mode=None)
    assert O.rel_l2(img, golden["small_even/numpy_chain_16x16"]) <= TOL
    out, mean, std = O.knee_chain_fastmri(k, m, (16, 16))
    assert O.rel_l2(out, golden["small_even/fastmri_chain_16x16"]) <= 5e-6
    np.testing.assert_allclose([mean, std], golden["small_even/fastmri_mean_std"], rtol=2e-6)


def test_small_odd(golden):
    k = golden["small_odd/kspace"]
    assert O.rel_l2(O.ifft2c(k), golden["small_odd/ifft2c"]) <= TOL
    assert O.rel_l2(O.fft2c(k), golden["small_odd/fft2c"]) <= TOL
    ri = O.to_real_view(k)
    assert O.rel_l2(O.ifft2c_new(ri), golden["small_odd/ifft2c_new"]) <= TOL
    assert O.rel_l2(O.fft2c_new(ri), golden["small_odd/fft2c_new"]) <= TOL
    assert O.rel_l2(O.rss_complex_ri(O.ifft2c_new(ri), dim=1), golden["small_odd/rss_complex"]) <= TOL
    assert O.rel_l2(O.rss(np.abs(k), dim=1), golden["small_odd/rss_real"]) <= TOL
    assert O.rel_l2(O.ifft2c_single(k[1, 2]), golden["small_odd/ifft2c_single"]) <= TOL
    assert O.rel_l2(O.ifftnd(k[0].copy(), [1, 2]), golden["small_odd/ifftnd"]) <= TOL
    # the three centred-iFFT spellings agree on odd sizes as well
    a = O.ifft2c(k)
    b = O.ifft2c_new(ri)
    assert O.rel_l2(b[..., 0] + 1j * b[..., 1], a) <= TOL


@pytest.mark.parametrize("name,gen", [("knee_gauss", synth.gaussian_kspace), ("knee_phantom", synth.phantom_kspace)])
def test_knee_config0(golden, manifest, name, gen):
    case = manifest["cases"][name]
    k = gen(tuple(case["shape"]), case["seed"])
    m = synth.knee_mask()
    np.testing.assert_array_equal(np.flatnonzero(m), golden["knee/mask_368_4x_008_idx"])
    assert int(m.sum()) == 114
    img, _, _ = O.knee_chain_numpy(k, m, synth.CROP, normalize_mode=None)
    assert img.shape == (320, 320) and img.dtype == np.float32
    assert O.rel_l2(img, golden[f"{name}/numpy_chain"]) <= TOL
    out, mean, std = O.knee_chain_fastmri(k, m, synth.CROP)
    assert O.rel_l2(out, golden[f"{name}/fastmri_chain"]) <= 5e-6
    np.testing.assert_allclose([mean, std], golden[f"{name}/fastmri_mean_std"], rtol=2e-6)
    # the numpy chain + normalize_instance and the fastMRI chain agree well inside the bar
    n_out, n_mean, n_std = O.normalize_instance(img)
    assert O.rel_l2(n_out, golden[f"{name}/fastmri_chain"]) <= 5e-6


def test_knee_nomask(golden):
    k = synth.gaussian_kspace(synth.KNEE_SHAPE, 0)
    img, _, _ = O.knee_chain_numpy(k, None, synth.CROP, normalize_mode=None)
    assert O.rel_l2(img[::4, ::4], golden["knee_gauss/numpy_chain_nomask_sub4"]) <= TOL


def test_single_coil(golden):
    a = O.ifft2c_single(synth.gaussian_kspace((640, 368), 1))
    assert O.rel_l2(a[::2], golden["single_coil/ifft2c_single_rows_even"]) <= TOL
    b = O.ifft2c_single(synth.gaussian_kspace((640, 372), 2))
    assert O.rel_l2(b[::5, ::3], golden["single_coil/ifft2c_single_372_sub"]) <= TOL
    with pytest.raises(ValueError):
        O.ifft2c_single(np.zeros((2, 8, 8), np.complex64))


def test_prostate_small(golden):
    assert golden["prostate_small/get_padding"][0] == 5.5
    assert O.padding_lr(32, 20) == (5, 6)
    k = golden["prostate_small/kspace"]
    for av in range(2):
        padded = O.zero_pad_pe(k[av], 5, 6)
        np.testing.assert_array_equal(padded.shape, golden[f"prostate_small/padded_shape_{av}"])
        assert O.rel_l2(O.create_coil_combined_im(padded), golden["prostate_small/coil_combined"][av]) <= TOL
    fin = O.t2_average_combine(k, (5, 6), (16, 16))
    assert fin.dtype == np.float64
    assert O.rel_l2(fin, golden["prostate_small/final_16x16"]) <= TOL


def test_prostate_config2_one_slice(golden, manifest):
    assert golden["prostate/get_padding_640_451"][0] == 94.5
    assert O.padding_lr(640, 450) == synth.PROSTATE_PAD == (94, 95)
    pm = synth.prostate_mask()
    np.testing.assert_array_equal(np.flatnonzero(pm), golden["prostate/mask_451_8x_004_idx"])
    case = manifest["cases"]["prostate_one_slice"]
    k = synth.gaussian_kspace(tuple(case["shape"]), case["seed"])
    fin = O.prostate_chain(k, pm, synth.PROSTATE_PAD)
    assert fin.shape == (1, 320, 320)
    assert O.rel_l2(fin, golden["prostate/one_slice_final"]) <= TOL


def test_normalize_instance_unbiased(golden):
    x = np.abs(synth.gaussian_kspace((320, 320), 3)).astype(np.float32) + 2.0
    out, mean, std = O.normalize_instance(x, eps=1e-11)
    np.testing.assert_allclose([mean, std], golden["norm/mean_std"], rtol=2e-6)
    assert O.rel_l2(out[::8, ::8], golden["norm/out_sub"]) <= 5e-6
    # N-1, not N: the population std differs by ~5e-6 relative, half the parity budget
    assert abs(float(std) - float(x.std(ddof=0))) / float(std) > 3e-6


def test_crop_indices_exact(golden):
    assert (O.crop_start(640, 320), O.crop_start(640, 320) + 320) == tuple(golden["index/crop_rows_640_320"])
    assert (O.crop_start(368, 320), O.crop_start(368, 320) + 320) == tuple(golden["index/crop_cols_368_320"])
    probe = np.arange(640 * 368, dtype=np.float32).reshape(640, 368)
    c = O.center_crop_or_pad(probe, 320, 320)
    np.testing.assert_array_equal(c[[0, 0, -1, -1], [0, -1, 0, -1]], golden["index/crop_probe_corners"])
    np.testing.assert_array_equal(c, O.center_crop(probe, (320, 320)))
    np.testing.assert_array_equal(c[None], O.center_crop_im(probe[None], [320, 320]))
    with pytest.raises(ValueError):
        O.center_crop(probe, (641, 320))
    for n, o in [(451, 320), (640, 320), (372, 320), (23, 7), (24, 7), (23, 8)]:
        x = np.arange(n * n, dtype=np.float64).reshape(1, n, n)
        np.testing.assert_array_equal(O.center_crop_im(x, [o, o]), O.center_crop(x, (o, o)))


def test_mask_rule_frozen():
    m = O.equispaced_mask(368, 4, 0.08)
    np.testing.assert_array_equal(m, synth.equispaced_mask(368, 4, 0.08))
    idx = np.flatnonzero(m)
    assert idx[0] == 0 and idx[-1] == 364
    assert set(range(170, 199)) <= set(idx.tolist()) and 169 not in idx and 199 not in idx
    assert int(O.equispaced_mask(368, 8, 0.04).sum()) == 60
    assert int(O.equispaced_mask(451, 8, 0.04).sum()) == len(np.flatnonzero(synth.prostate_mask()))


def test_frozen_mask_index_lists(manifest):
    """the mask generator is builder-defined, so its index lists are frozen (manifest "masks")."""
    assert set(manifest["masks"]) == {synth.mask_name(*spec) for spec in synth.FROZEN_MASKS}
    for spec in synth.FROZEN_MASKS:
        want = manifest["masks"][synth.mask_name(*spec)]
        for gen in (synth.equispaced_mask, O.equispaced_mask):
            m = gen(*spec)
            assert m.dtype == np.float32 and set(np.unique(m).tolist()) <= {0.0, 1.0}
            assert np.flatnonzero(m).tolist() == want, spec
    assert len(manifest["masks"]["equispaced(368,4,0.08,0)"]) == 114
    assert len(manifest["masks"]["equispaced(368,8,0.04,0)"]) == 60


@pytest.mark.parametrize("tag,spec", [("8x", (368, 8, 0.04, 0)), ("4x_off1", (368, 4, 0.08, 1)), ("4x_off3", (368, 4, 0.08, 3))])
def test_knee_other_masks(golden, tag, spec):
    k = synth.gaussian_kspace(synth.KNEE_SHAPE, 0)
    m = synth.equispaced_mask(*spec)
    img, _, _ = O.knee_chain_numpy(k, m, synth.CROP, normalize_mode=None)
    assert O.rel_l2(img[::2, ::2], golden[f"knee_gauss/numpy_chain_{tag}_sub2"]) <= TOL
    _, mean, std = O.knee_chain_fastmri(k, m, synth.CROP)
    np.testing.assert_allclose([mean, std], golden[f"knee_gauss/fastmri_mean_std_{tag}"], rtol=2e-6)


@pytest.mark.parametrize("s_idx", [0, 29])
def test_prostate_volume_slices(golden, s_idx):
    """two slices of the full configs[2] volume, built block by block (synth.prostate_volume_block)"""
    k = np.stack([synth.prostate_volume_block(a, s_idx) for a in range(3)])[:, None]
    fin = O.prostate_chain(k, synth.prostate_mask(), synth.PROSTATE_PAD)
    assert O.rel_l2(fin[0, ::2, ::2], golden[f"prostate/volume_slice{s_idx}_sub2"]) <= TOL


def test_analytic_cases():
    # delta at the k-space centre <-> constant image 1/sqrt(HW); Parseval for the ortho pair
    h, w = 640, 368
    k = np.zeros((h, w), np.complex64)
    k[h // 2, w // 2] = 1.0
    np.testing.assert_allclose(O.ifft2c_single(k), 1.0 / np.sqrt(h * w), rtol=1e-6)
    x = synth.gaussian_kspace((3, 30, 23), 5)
    assert abs(np.linalg.norm(O.ifft2c(x)) / np.linalg.norm(x) - 1) < 1e-6
    assert O.rel_l2(O.fft2c(O.ifft2c(x)), x) <= TOL
