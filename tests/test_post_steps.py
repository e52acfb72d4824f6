"""The per-slice steps after the reconstruction (REF/src/preprocess/mri_preprocess.py:182-191,216-233).

CPU: the oracle's restatement against the golden vectors frozen from the live reference (oracle/make_golden_post.py) and --
where /root/reference exists -- against the reference itself.  GPU: the CUDA kernels through the Python twins / the C ABI
against both.  Tolerances: percentiles and resized masks bit-exact; float images 2e-6 of the image's dynamic range (the
reference's torch CPU interpolation and numpy's pairwise float32 means round differently in the last place)."""
import json
import os

import numpy as np
import pytest
import torch

from mri_acl_imagesegmentation_adsp_b200 import synth
from oracle import recon_oracle as O
from oracle import ref_shim

HERE = os.path.dirname(os.path.abspath(__file__))
CLIP = (1.0, 99.5)


@pytest.fixture(scope="module")
def post_golden():
    return np.load(os.path.join(HERE, "golden", "post_vectors.npz")), json.load(open(os.path.join(HERE, "golden", "post_manifest.json")))


def _close(a, b, scale):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max()) <= 2e-6 * scale


CASES = [c[0] for c in synth.POST_CASES]


@pytest.mark.parametrize("name", CASES)
def test_oracle_post_chain_vs_golden(post_golden, name):
    g, man = post_golden
    c = man["cases"][name]
    img, mk, out = synth.post_case_inputs(name)
    z, p01, mr, (lo, hi) = O.post_chain(img, mk, out, CLIP)
    sub = c["float_subsample"]
    assert np.array_equal(np.array([lo, hi], np.float32), g[f"{name}/lo_hi"])
    assert O.percentile_f32(img, CLIP[0]) == g[f"{name}/lo_hi"][0] and O.percentile_f32(img, CLIP[1]) == g[f"{name}/lo_hi"][1]
    assert np.array_equal(mr, g[f"{name}/mask_r"])
    img_r = O.resize_bilinear(O.percentile_clip(img, *CLIP)[0], out)
    assert _close(img_r[::sub, ::sub], g[f"{name}/img_r"], float(img.max()))
    assert _close(z[::sub, ::sub], g[f"{name}/img_z"], max(1.0, float(np.abs(g[f"{name}/img_z"]).max())))
    assert _close(p01[::sub, ::sub], g[f"{name}/img_01"], 4.0)        # (tiny-mask case: values far outside [0, 1])


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")
def test_oracle_post_steps_vs_live_reference():
    pre = ref_shim.knee_preprocessor_cls()
    rng = np.random.default_rng(9)
    for shape, out in (((64, 48), (32, 32)), ((33, 47), (50, 20)), ((320, 320), (320, 320))):
        img = np.abs(rng.standard_normal(shape)).astype(np.float32) * 3
        mk = (img > 1.0).astype(np.uint8)
        for q in (0.0, 1.0, 37.3, 50.0, 99.5, 100.0):
            assert O.percentile_f32(img, q) == np.percentile(img, q)
        want = pre._percentile_clip(img, *CLIP)
        assert np.array_equal(O.percentile_clip(img, *CLIP)[0], want)
        assert _close(O.resize_bilinear(want, out), pre._resize_np(want, out), 10.0)
        assert np.array_equal(O.resize_mask(mk, out), (pre._resize_np(mk.astype(np.float32), out) > 0.5).astype(np.uint8))
        r = pre._resize_np(want, out)
        mr = O.resize_mask(mk, out)
        assert np.array_equal(O.zscore_in_mask(r, mr), pre._zscore_in_mask(r, mr))
        assert np.array_equal(O.preview_01(r, mr), pre._preview_01(r, mr))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_post_chain_vs_golden_and_oracle(post_golden, name):
    from mri_acl_imagesegmentation_adsp_b200.preprocess.mri_preprocess import MRIKneePreprocessor
    g, man = post_golden
    c = man["cases"][name]
    sub = c["float_subsample"]
    img, mk, out = synth.post_case_inputs(name)
    pre = MRIKneePreprocessor(out_size=out, clip_percentiles=CLIP)
    r = pre.clip_resize_zscore(torch.from_numpy(img[None]).cuda(), torch.from_numpy(mk[None]).cuda())
    assert np.array_equal(r["clip"][0].cpu().numpy(), g[f"{name}/lo_hi"])                      # percentiles: bit-exact
    assert np.array_equal(r["mask"][0].cpu().numpy(), g[f"{name}/mask_r"])                     # resized mask: bit-exact
    z, p01 = r["img_z"][0].cpu().numpy(), r["img_01"][0].cpu().numpy()
    assert _close(z[::sub, ::sub], g[f"{name}/img_z"], max(1.0, float(np.abs(g[f"{name}/img_z"]).max())))
    assert _close(p01[::sub, ::sub], g[f"{name}/img_01"], 4.0)
    wz, wp, wm, _ = O.post_chain(img, mk, out, CLIP)
    assert _close(z, wz, max(1.0, float(np.abs(wz).max()))) and _close(p01, wp, 4.0) and np.array_equal(r["mask"][0].cpu().numpy(), wm)
    # the twins of the reference's static methods, one step at a time (numpy in -> numpy out)
    clipped = MRIKneePreprocessor._percentile_clip(img, *CLIP)
    assert isinstance(clipped, np.ndarray) and np.array_equal(clipped, O.percentile_clip(img, *CLIP)[0])
    step = c["clipped_row_step"]
    assert np.array_equal(clipped[::step, ::sub], g[f"{name}/clipped"])
    img_r = MRIKneePreprocessor._resize_np(clipped, out)
    assert _close(img_r[::sub, ::sub], g[f"{name}/img_r"], float(img.max()))
    assert _close(MRIKneePreprocessor._zscore_in_mask(img_r, wm), O.zscore_in_mask(img_r, wm), max(1.0, float(np.abs(wz).max())))
    assert _close(MRIKneePreprocessor._preview_01(img_r, wm), O.preview_01(img_r, wm), 4.0)


@pytest.mark.gpu
def test_gpu_percentiles_exact_on_hard_inputs():
    """ties, negative values, denormals, extreme percentiles, a batch of different images in one call."""
    from mri_acl_imagesegmentation_adsp_b200.adapters import recon_cabi
    lib = recon_cabi.library()
    rng = np.random.default_rng(3)
    imgs = np.stack([rng.standard_normal(5000).astype(np.float32) * 10,
                     np.round(rng.standard_normal(5000) * 3).astype(np.float32),            # heavy ties, +-0
                     (rng.standard_normal(5000) * 1e-41).astype(np.float32),               # denormals
                     np.abs(rng.standard_normal(5000)).astype(np.float32) ** 8])           # huge dynamic range
    t = torch.from_numpy(imgs).cuda()
    out = torch.empty_like(t)
    lh = torch.empty((4, 2), dtype=torch.float32, device="cuda")
    for pmin, pmax in ((1.0, 99.5), (0.0, 100.0), (37.3, 37.4), (49.99, 50.01)):
        lib.percentile_clip(t.data_ptr(), out.data_ptr(), lh.data_ptr(), 4, 5000, pmin, pmax, 0)
        torch.cuda.synchronize()
        for b in range(4):
            lo, hi = np.percentile(imgs[b], pmin), np.percentile(imgs[b], pmax)
            assert lh[b, 0].item() == lo and lh[b, 1].item() == hi, (b, pmin, pmax)
            assert np.array_equal(out[b].cpu().numpy(), np.clip(imgs[b], lo, hi))
    with pytest.raises(ValueError):
        lib.percentile_clip(t.data_ptr(), out.data_ptr(), lh.data_ptr(), 4, 5000, 60.0, 40.0, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("B,n", [(3, 5003), (3, 102400), (40, 5003), (40, 10240), (80, 4099), (80, 4096)])
def test_gpu_one_image_kernels_for_every_cluster_size(B, n):
    """the percentile and z-score kernels run as clusters of 4 / 2 / 1 CTAs per image depending on the batch (B x CL within
    one wave of SMs): bit-exact percentiles and clipped images, z-score / preview / statistics against the oracle, for
    vector-aligned and odd image sizes (ragged shares per CTA) -- and the same answers whatever the cluster size, checked
    by running each image again in a batch of another size."""
    from mri_acl_imagesegmentation_adsp_b200.adapters import recon_cabi
    lib = recon_cabi.library()
    rng = np.random.default_rng(B * 7 + n)
    imgs = (rng.standard_normal((B, n)).astype(np.float32) ** 3) * np.float32(4.0)
    imgs[1] = np.round(imgs[1])                                      # ties
    masks = (rng.uniform(size=(B, n)) < 0.6).astype(np.uint8)
    masks[2] = 0                                                     # empty mask: statistics of the whole image
    t, mk = torch.from_numpy(imgs).cuda(), torch.from_numpy(masks).cuda()
    out, lh = torch.empty_like(t), torch.empty((B, 2), dtype=torch.float32, device="cuda")
    z, q01, st = torch.empty_like(t), torch.empty_like(t), torch.empty((B, 6), dtype=torch.float32, device="cuda")
    lib.percentile_clip(t.data_ptr(), out.data_ptr(), lh.data_ptr(), B, n, CLIP[0], CLIP[1], 0)
    lib.zscore_preview(t.data_ptr(), mk.data_ptr(), z.data_ptr(), q01.data_ptr(), st.data_ptr(), B, n, 0)
    torch.cuda.synchronize()
    for b in sorted({0, 1, 2, B // 2, B - 1}):
        lo, hi = np.percentile(imgs[b], CLIP[0]), np.percentile(imgs[b], CLIP[1])
        assert lh[b, 0].item() == lo and lh[b, 1].item() == hi, b
        assert np.array_equal(out[b].cpu().numpy(), np.clip(imgs[b], lo, hi))
        zr, pr = O.zscore_in_mask(imgs[b], masks[b]), O.preview_01(imgs[b], masks[b])
        scale = float(np.abs(zr).max())
        assert np.abs(z[b].cpu().numpy() - zr).max() <= 2e-6 * max(1.0, scale)
        assert np.abs(q01[b].cpu().numpy() - pr).max() <= 2e-6 * max(1.0, float(np.abs(pr).max()))
    # the same images in a batch that takes another cluster size: identical percentiles, statistics within rounding
    B2 = 2 if B > 3 else 80
    idx = np.arange(B2) % B
    t2, mk2 = t[torch.from_numpy(idx).cuda()].contiguous(), mk[torch.from_numpy(idx).cuda()].contiguous()
    lh2, st2 = torch.empty((B2, 2), dtype=torch.float32, device="cuda"), torch.empty((B2, 6), dtype=torch.float32, device="cuda")
    z2 = torch.empty_like(t2)
    lib.percentile_clip(t2.data_ptr(), 0, lh2.data_ptr(), B2, n, CLIP[0], CLIP[1], 0)
    lib.zscore_preview(t2.data_ptr(), mk2.data_ptr(), z2.data_ptr(), 0, st2.data_ptr(), B2, n, 0)
    torch.cuda.synchronize()
    assert torch.equal(lh2, lh[torch.from_numpy(idx).cuda()])
    torch.testing.assert_close(st2, st[torch.from_numpy(idx).cuda()], rtol=1e-6, atol=0)


@pytest.mark.gpu
def test_gpu_preprocess_records_contract():
    """the k-space branch of preprocess_records on the device: slice_keep band, (S,1,H,W) float32 tensor, previews, masks."""
    from mri_acl_imagesegmentation_adsp_b200.preprocess.mri_preprocess import MRIKneePreprocessor
    ks = [synth.phantom_kspace((1, 64, 48), 40 + i)[0] for i in range(10)]
    recs = [{"kspace": k, "meta": {"slice_idx": 100 + i}} for i, k in enumerate(ks)]
    fn = lambda im: synth.body_mask_standin(im, 0.3)
    pre = MRIKneePreprocessor(out_size=(32, 32), body_mask_fn=fn)
    out = pre.preprocess_records(recs)
    assert out["tensor"].shape == (4, 1, 32, 32) and out["tensor"].dtype == torch.float32 and out["tensor"].device.type == "cpu"
    assert out["preview"].shape == (4, 32, 32) and out["mask"].shape == (4, 32, 32) and out["mask"].dtype == np.uint8
    assert out["indices"] == [103, 104, 105, 106] and out["sources"] == ["kspace"] * 4
    for j, i in enumerate(range(3, 7)):
        img = O.ifft2c_single(ks[i])
        clipped = O.percentile_clip(img, *CLIP)[0]
        z, p01, mr, _ = O.post_chain(img, fn(clipped), (32, 32), CLIP)
        assert np.array_equal(out["mask"][j], mr)
        assert _close(out["tensor"][j, 0].numpy(), z, max(1.0, float(np.abs(z).max()))) and _close(out["preview"][j], p01, 4.0)
    with pytest.raises(ValueError):
        pre.preprocess_records([])
    with pytest.raises(ValueError):
        MRIKneePreprocessor(use_n4=True)


def _dataset_getitem_inputs(vol, k, imagenet):
    """what KneeNPZ2DSlices.__getitem__ (REF/src/dataio/datasets.py:86-131) hands the network for every slice of one volume
    (no augmentation), restated with the same numpy / torch calls."""
    S = vol.shape[0]
    mean = torch.tensor((0.485, 0.456, 0.406)).view(-1, 1, 1)
    std = torch.tensor((0.229, 0.224, 0.225)).view(-1, 1, 1)
    outs = []
    for s in range(S):
        if k == 1:
            x = vol[s]
        else:
            half = k // 2
            idxs = [min(max(s + d, 0), S - 1) for d in range(-half, half + 1)]
            x = np.concatenate([vol[j] for j in idxs], axis=0)
        t = torch.from_numpy(x.copy()).float()
        if imagenet and t.shape[0] == 1:
            t = t.repeat(3, 1, 1)
        if imagenet:
            t = (t - mean) / std
        outs.append(t.contiguous())
    return torch.stack(outs)


@pytest.mark.gpu
def test_gpu_stack_2p5d_matches_the_dataset_rule():
    from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import stack_2p5d
    rng = np.random.default_rng(12)
    vol = rng.standard_normal((7, 1, 20, 24)).astype(np.float32)
    for k, imagenet in ((1, False), (3, False), (5, False), (1, True), (3, True)):
        got = stack_2p5d(vol, k, imagenet_norm=imagenet)
        want = _dataset_getitem_inputs(vol, k, imagenet).numpy()
        assert got.shape == want.shape and np.array_equal(got, want), (k, imagenet)          # bit-exact, normalisation included
    dev = stack_2p5d(torch.from_numpy(vol).cuda(), 3)
    assert dev.is_cuda and dev.shape == (7, 3, 20, 24)
    with pytest.raises(ValueError):
        stack_2p5d(vol, 2)
    with pytest.raises(ValueError):
        stack_2p5d(vol, 5, imagenet_norm=True)           # five channels against three ImageNet means
