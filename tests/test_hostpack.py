"""Host side of the end-to-end path: mriacl_pack_columns_host (plain C++ threads, no CUDA) against numpy indexing,
and -- on the GPU -- the packed-column transfer mode of HostPipeline against the direct one (bit-identical images)."""
import numpy as np
import pytest
import torch

from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.adapters import recon_cabi


def _lib():
    return recon_cabi.ReconLibrary(recon_cabi.DEFAULT_LIBRARY)


@pytest.mark.parametrize("threads", [1, 3, 0])
def test_pack_columns_host_matches_numpy(threads):
    lib = _lib()
    rng = np.random.default_rng(5)
    for rows, W, mask in ((640 * 3, 368, synth.knee_mask()), (1000, 451, synth.prostate_mask()), (7, 23, None),
                          (129, 64, (rng.uniform(size=64) < 0.4).astype(np.float32) * 0.5),     # weights: value != 0 selects
                          (0, 16, np.ones(16, np.float32)), (5, 9, np.zeros(9, np.float32))):
        src = synth.gaussian_kspace((max(rows, 1), W), 11)[:rows]
        idx = np.arange(W) if mask is None else np.nonzero(mask)[0]
        dst = np.full((rows, max(1, len(idx))), -7 - 7j, np.complex64)
        n = lib.pack_columns_host(src.ctypes.data, dst.ctypes.data, rows, W, mask, threads)
        assert n == len(idx)
        if len(idx):
            got = dst.reshape(-1)[:rows * n].reshape(rows, n)       # the packed buffer is dense [rows][n_act]
            assert np.array_equal(got.view(np.uint64), np.ascontiguousarray(src[:, idx]).view(np.uint64))   # bit-exact


def test_pack_columns_host_rejects_bad_arguments():
    lib = _lib()
    a = np.zeros(4, np.complex64)
    with pytest.raises(ValueError):
        lib.pack_columns_host(a.ctypes.data, a.ctypes.data, 1, 0, None, 1)
    with pytest.raises(ValueError):
        lib.pack_columns_host(a.ctypes.data, a.ctypes.data, -1, 4, None, 1)
    with pytest.raises(ValueError):
        lib.pack_columns_host(0, a.ctypes.data, 1, 4, None, 1)


@pytest.mark.gpu
def test_packed_pipeline_is_bit_identical():
    from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import zero_filled_rss
    from mri_acl_imagesegmentation_adsp_b200.recon.pipeline import HostPipeline
    from oracle import recon_oracle as O
    k_np = synth.gaussian_kspace((11,) + synth.KNEE_SHAPE, 21)
    k_host = torch.from_numpy(k_np)                      # pageable on purpose: the packed mode does not need pinned input
    outs = {}
    for m in (synth.knee_mask(), synth.equispaced_mask(368, 8, 0.04), synth.equispaced_mask(368, 4, 0.08, 3)):
        dev_img, dev_mean, dev_std = zero_filled_rss(torch.from_numpy(k_np).cuda(), m, synth.CROP, "instance")
        for pack, every in ((False, 0), (True, 0), ("auto", 0), (True, 2), (True, 3)):
            # (True, k): mixed transfer, sub-batches 0, k, 2k, ... go across full width, the others packed
            out = torch.empty((11,) + synth.CROP, dtype=torch.float32).pin_memory()
            ms = torch.empty((11, 2), dtype=torch.float32).pin_memory()
            pipe = HostPipeline(synth.KNEE_SHAPE, synth.CROP, "instance", 0.0, sub_batch=2 if every else 4, n_streams=2, pack=pack,
                                direct_every=every)
            pipe(k_host.pin_memory() if pack is False or every else k_host, m, out, ms)
            if every:
                n_direct = len(range(0, 6, every))                  # six sub-batches of two slices (the last holds one)
                n_act = int(np.count_nonzero(m))
                assert pipe.h2d_bytes == (2 * n_direct * 368 + (11 - 2 * n_direct) * n_act) * 15 * 640 * 8
            torch.cuda.synchronize()
            assert torch.equal(out, dev_img.cpu()), f"pack={pack}"
            assert torch.equal(ms[:, 0], dev_mean.cpu()) and torch.equal(ms[:, 1], dev_std.cpu())
            if pack is True and not every:
                assert pipe.h2d_bytes == 11 * 15 * 640 * int(np.count_nonzero(m)) * 8
            if pack == "auto":
                assert pipe.calibration is not None and pipe.pack in (True, False) and "chosen" in pipe.calibration
        want, _, _ = O.knee_chain_numpy(k_np[10], m, synth.CROP, "instance")
        assert O.rel_l2(out[10].numpy(), want) <= 1e-5


@pytest.mark.gpu
def test_packed_kspace_other_plans():
    """packed columns through the 640-wide (prostate) plan with averages / pad / flip, a 372-wide plan and a pruned-generic
    width; shapes without the 640-row column pass refuse the flag."""
    from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import zero_filled_rss
    kp = synth.gaussian_kspace((2, 3, 4, 640, 451), 31)       # (A, S, C, RO, PE)
    pm = synth.prostate_mask()
    idx = np.nonzero(pm)[0]
    full, _, _ = zero_filled_rss(torch.from_numpy(kp).cuda(), pm, synth.CROP, None, flip_rows=True, average_axis=0, pad=(94, 95))
    pk, _, _ = zero_filled_rss(torch.from_numpy(np.ascontiguousarray(kp[..., idx])).cuda(), pm, synth.CROP, None, flip_rows=True,
                               average_axis=0, pad=(94, 95), packed=True)
    assert torch.equal(full, pk)
    for W in (372, 320):
        k = synth.gaussian_kspace((3, 5, 640, W), 32)
        m = synth.equispaced_mask(W, 4, 0.08)
        a, _, _ = zero_filled_rss(torch.from_numpy(k).cuda(), m, (320, 320), "instance")
        b, _, _ = zero_filled_rss(torch.from_numpy(np.ascontiguousarray(k[..., np.nonzero(m)[0]])).cuda(), m, (320, 320), "instance",
                                  packed=True)
        assert torch.equal(a, b)
    k = synth.gaussian_kspace((1, 2, 64, 48), 33)
    m = synth.equispaced_mask(48, 4, 0.08)
    with pytest.raises(ValueError):
        zero_filled_rss(torch.from_numpy(np.ascontiguousarray(k[..., np.nonzero(m)[0]])).cuda(), m, (32, 32), None, packed=True)
    with pytest.raises(ValueError):
        zero_filled_rss(torch.from_numpy(k).cuda(), m, (32, 32), None, packed=True)      # width != number of sampled columns
