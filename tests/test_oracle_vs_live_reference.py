"""oracle/recon_oracle.py against the LIVE reference functions (build container only;
skipped where /root/reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from mri_acl_imagesegmentation_adsp_b200 import synth
from oracle import recon_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")


@pytest.mark.parametrize("shape", [(2, 16, 12), (3, 15, 11), (1, 30, 23), (2, 2, 8, 9)])
def test_transforms_bitwise(shape):
    ks = ref_shim.kspace_utils()
    x = synth.gaussian_kspace(shape, 21)
    np.testing.assert_array_equal(O.ifft2c(x), ks.ifft2c(x))
    np.testing.assert_array_equal(O.fft2c(x), ks.fft2c(x))
    np.testing.assert_array_equal(O.complex_abs(x), ks.complex_abs(x))
    for oh, ow in [(8, 8), (40, 5), (shape[-2], shape[-1]), (3, 50)]:
        np.testing.assert_array_equal(O.center_crop_or_pad(x, oh, ow), ks.center_crop_or_pad(x, oh, ow))
    pre = ref_shim.knee_preprocessor_cls()
    flat = x.reshape((-1,) + x.shape[-2:])[0]
    np.testing.assert_array_equal(O.ifft2c_single(flat), pre.ifft2c_single(flat))
    with pytest.raises(ValueError):
        pre.ifft2c_single(x)


@pytest.mark.parametrize("shape", [(4, 16, 12), (3, 15, 11), (5, 31, 24)])
def test_fastmri_twins(shape):
    dl = ref_shim.fastmri_dl()
    x = synth.gaussian_kspace(shape, 22)
    t = dl.transforms.to_tensor(x)
    ri = O.to_real_view(x)
    np.testing.assert_array_equal(ri, t.numpy())
    assert O.rel_l2(O.ifft2c_new(ri), dl.fftc.ifft2c_new(t).numpy()) <= 1e-6
    assert O.rel_l2(O.fft2c_new(ri), dl.fftc.fft2c_new(t).numpy()) <= 1e-6
    img = dl.fftc.ifft2c_new(t)
    assert O.rel_l2(O.rss_complex_ri(img.numpy(), 0), dl.coil_combine.rss_complex(img, 0).numpy()) <= 1e-6
    assert O.rel_l2(O.complex_abs_ri(img.numpy()), dl.math_fn.complex_abs(img).numpy()) <= 1e-6
    mag = dl.math_fn.complex_abs(img)
    assert O.rel_l2(O.rss(mag.numpy(), 0), dl.coil_combine.rss(mag, 0).numpy()) <= 1e-6
    crop = (shape[1] // 2, shape[2] // 2)
    np.testing.assert_array_equal(O.center_crop(mag.numpy(), crop), dl.transforms.center_crop(mag, crop).numpy())
    with pytest.raises(ValueError):
        dl.transforms.center_crop(mag, (shape[1] + 1, 2))
    with pytest.raises(ValueError):
        dl.fftc.ifft2c_new(torch.zeros(4, 4, 3))
    comb = dl.coil_combine.rss_complex(img, 0)
    o, mu, sd = dl.transforms.normalize_instance(comb, eps=1e-11)
    oo, omu, osd = O.normalize_instance(comb.numpy(), eps=1e-11)
    assert O.rel_l2(oo, o.numpy()) <= 5e-6
    np.testing.assert_allclose([omu, osd], [mu.item(), sd.item()], rtol=2e-6)


def test_prostate_twins():
    pr = ref_shim.prostate()
    k = synth.gaussian_kspace((2, 3, 4, 24, 17), 23)
    hdr = ref_shim.synthetic_header(24, 16)
    l, r = O.padding_lr(24, 16)
    ims = []
    for a in range(2):
        padded = pr.mri_data.zero_pad_kspace_hdr(hdr, k[a])
        np.testing.assert_array_equal(padded, O.zero_pad_pe(k[a], l, r))
        ref_im = pr.t2.create_coil_combined_im(padded)
        np.testing.assert_array_equal(ref_im, O.create_coil_combined_im(padded))
        ims.append(ref_im)
    ref_fin = pr.utils.center_crop_im(np.mean(np.stack(ims), axis=0), [12, 12])
    np.testing.assert_array_equal(ref_fin, O.t2_average_combine(k, (l, r), (12, 12)))
    for enc_x, max_pe in [(640, 450), (640, 447), (32, 20), (32, 19), (64, 31)]:
        p = pr.mri_data.get_padding(ref_shim.synthetic_header(enc_x, max_pe))
        lr = O.padding_lr(enc_x, max_pe)
        assert lr[0] + lr[1] == enc_x - (max_pe + 1)
        assert lr == ((int(np.floor(p)), int(np.ceil(p))) if p % 2 != 0 else (int(p), int(p)))


def test_reference_signature_twins():
    """ifftnd (default axes=[-1], single axes, None), flip_im (with its first-axis quirk), center_crop_im, numpy rss,
    complex_center_crop, center_crop_to_smallest: the oracle restatements against the live functions."""
    pr = ref_shim.prostate()
    dl = ref_shim.fastmri_dl()
    x = synth.gaussian_kspace((3, 10, 7), 24)
    np.testing.assert_array_equal(O.ifftnd(x.copy()), pr.utils.ifftnd(x.copy()))
    for axes in ([0], [1], [0, 2], [1, 2], None):
        np.testing.assert_array_equal(O.ifftnd(x.copy(), axes), pr.utils.ifftnd(x.copy(), axes))
    x128 = x.astype(np.complex128)
    assert pr.utils.ifftnd(x128.copy(), [1, 2]).dtype == np.complex128 == O.ifftnd(x128.copy(), [1, 2]).dtype
    vol = np.abs(x).astype(np.float64)
    for ax in (0, 1):
        np.testing.assert_array_equal(O.flip_im(vol[:, :3].copy(), ax), pr.utils.flip_im(vol[:, :3].copy(), ax))
    np.testing.assert_array_equal(O.center_crop_im(vol, [4, 6]), pr.utils.center_crop_im(vol, [4, 6]))
    for ax in (-1, 0, 1):
        np.testing.assert_array_equal(O.rss_np(x, ax), pr.t2.rss(x, ax))
    assert pr.t2.rss(x, 0).dtype == np.float32 and pr.t2.rss(x128, 0).dtype == np.float64
    t = dl.transforms.to_tensor(x)
    np.testing.assert_array_equal(O.complex_center_crop_ri(t.numpy(), (6, 3)), dl.transforms.complex_center_crop(t, (6, 3)).numpy())
    with pytest.raises(ValueError):
        dl.transforms.complex_center_crop(t, (11, 3))
    a, b = torch.from_numpy(vol[:, :8, :]), torch.from_numpy(vol[:, :, :5])
    ra, rb = dl.transforms.center_crop_to_smallest(a, b)
    oa, ob = O.center_crop_to_smallest(a.numpy(), b.numpy())
    np.testing.assert_array_equal(ra.numpy(), oa)
    np.testing.assert_array_equal(rb.numpy(), ob)
