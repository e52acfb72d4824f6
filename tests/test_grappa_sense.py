"""GRAPPA and the SENSE-style combine (row 4).

CPU: the host twin's geometry extraction and weight solve, and the oracle's restatement, against golden vectors frozen from
the live vendored class (oracle/make_golden_grappa.py) and -- where /root/reference exists -- against the class itself
(bit-equal: both are the same numpy calls).  GPU: the weight application kernel and the combine through the twins."""
import json
import os

import numpy as np
import pytest
import torch

from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.prostate.grappa import Grappa
from oracle import recon_oracle as O
from oracle import ref_shim

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = [c[0] for c in synth.GRAPPA_CASES]


@pytest.fixture(scope="module")
def gg():
    return np.load(os.path.join(HERE, "golden", "grappa_vectors.npz")), json.load(open(os.path.join(HERE, "golden", "grappa_manifest.json")))


@pytest.mark.parametrize("name", NAMES)
def test_host_twin_and_oracle_vs_golden(gg, name):
    g, man = gg
    sub = man["cases"][name]["ro_subsample"]
    k, calib = synth.grappa_case_inputs(name)
    tw = Grappa(k.copy(), kernel_size=(5, 5), coil_axis=1)
    keys = [int(i) for i in tw.kernel_var_dict["patch_indices"]]
    assert keys == g[f"{name}/patch_indices"].tolist()                                     # geometry keys: exact
    assert [len(tw.kernel_var_dict["holes_x"][i]) for i in keys] == g[f"{name}/holes_per_geometry"].tolist()
    kp, P, valid, hx, hy = O.grappa_geometries(k, (5, 5), 1)
    assert valid.tolist() == keys
    for i in keys:                                                                         # twin == literal restatement
        assert np.array_equal(tw.kernel_var_dict["holes_x"][i], hx[i]) and np.array_equal(tw.kernel_var_dict["holes_y"][i], hy[i])
        assert np.array_equal(tw.kernel_var_dict["patches"][i][..., 0], P[i])
    w = tw.compute_weights(calib.copy())
    wo = O.grappa_weights(calib, P, valid, (5, 5), 1)
    for i in keys:
        assert O.rel_l2(w[i], g[f"{name}/weights_{i}"]) <= 1e-6 and np.array_equal(w[i], wo[i])
    assert O.rel_l2(O.grappa_apply(k, w, (5, 5), 1)[:, :, ::sub], g[f"{name}/filled"]) <= 1e-6


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")
def test_host_twin_vs_live_reference():
    RG = ref_shim.prostate().grappa.Grappa
    k, calib = synth.grappa_case_inputs("small_r2")
    k[5, :, 7] = 0                                   # an isolated data zero on a sampled line: its own geometries
    for coil_axis, kk, cc in ((1, k, calib), (-1, np.moveaxis(k, 1, -1).copy(), np.moveaxis(calib, 1, -1).copy())):
        a, b = RG(kk.copy(), (5, 5), coil_axis), Grappa(kk.copy(), (5, 5), coil_axis)
        ka, kb = a.kernel_var_dict, b.kernel_var_dict
        assert np.array_equal(ka["patch_indices"], kb["patch_indices"]) and np.array_equal(ka["patches"], kb["patches"])
        for i in ka["patch_indices"]:
            assert np.array_equal(ka["holes_x"][i], kb["holes_x"][i]) and np.array_equal(ka["holes_y"][i], kb["holes_y"][i])
        wa, wb = a.compute_weights(cc.copy()), b.compute_weights(cc.copy())
        assert all(np.array_equal(wa[i], wb[i]) for i in wa)
        assert np.array_equal(O.grappa_apply(kk, wa, (5, 5), coil_axis), a.apply_weights(kk.copy(), wa))
    full = synth.gaussian_kspace((8, 3, 6), 1)       # no holes: the reference keeps the k-space as "geometries"
    assert isinstance(Grappa(full.copy(), (5, 5), 1).kernel_var_dict, np.ndarray)
    assert isinstance(RG(full.copy(), (5, 5), 1).kernel_var_dict, np.ndarray)


def test_sense_oracle_vs_golden(gg):
    g, _ = gg
    img, sens = synth.gaussian_kspace((3, 5, 12, 10), 611), synth.gaussian_kspace((3, 5, 12, 10), 612)
    assert O.rel_l2(O.sense_combine(img, sens, True), g["sense/abs_sum"]) <= 1e-6


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_grappa_apply_vs_golden(gg, name):
    g, man = gg
    sub = man["cases"][name]["ro_subsample"]
    k, calib = synth.grappa_case_inputs(name)
    tw = Grappa(k.copy(), kernel_size=(5, 5), coil_axis=1)
    w = tw.compute_weights(calib.copy())
    out = tw.apply_weights(k.copy(), w)
    assert isinstance(out, np.ndarray) and out.shape == k.shape and out.dtype == np.complex64
    assert O.rel_l2(out[:, :, ::sub], g[f"{name}/filled"]) <= 1e-5
    keep = np.abs(k[:, 0, :]) > 0
    assert np.array_equal(out.transpose(0, 2, 1)[keep], k.transpose(0, 2, 1)[keep])         # sampled positions: untouched, bit-exact
    # torch CUDA in -> torch CUDA out, the input is not modified
    kt = torch.from_numpy(k).cuda()
    ot = tw.apply_weights(kt, w)
    assert ot.is_cuda and torch.equal(kt.cpu(), torch.from_numpy(k)) and np.array_equal(ot.cpu().numpy(), out)


@pytest.mark.gpu
def test_gpu_grappa_batch_on_file_layout():
    """(S, C, RO, PE) slices of one average, each with its own weights, filled in ONE launch on the file's axis order
    (kernel axes = (PE, RO), coil axis first), against the oracle slice by slice; coil-last layout as well."""
    rng = np.random.default_rng(7)
    S, C, RO, PE = 3, 6, 40, 33
    keep = np.zeros(PE, bool); keep[1::2] = True; keep[12:20] = True
    k = synth.gaussian_kspace((S, C, RO, PE), 71)
    k[..., ~keep] = 0
    calibs = [synth.grappa_case_inputs("small_r2")[1][:, :1].repeat(C, 1) * (1 + 0.1 * s) + 0.05 * synth.gaussian_kspace((12, C, 20), 80 + s)
              for s in range(S)]
    tw = Grappa(np.transpose(k[0], (2, 0, 1)).copy(), kernel_size=(5, 5), coil_axis=1)      # (PE, C, RO) as prostate_t2_recon.py:34
    ws = [tw.compute_weights(c.astype(np.complex64)) for c in calibs]
    out = tw.apply_weights_batch(k.copy(), ws, axes=(2, 1, 0))                               # x = PE (axis 2), y = RO (axis 1), coil = axis 0
    for s in range(S):
        want = O.grappa_apply(np.transpose(k[s], (2, 0, 1)), ws[s], (5, 5), 1)               # (PE, C, RO)
        assert O.rel_l2(np.transpose(out[s], (2, 0, 1)), want) <= 1e-5
    kl = np.ascontiguousarray(np.transpose(k, (0, 3, 2, 1)))                                 # (S, PE, RO, C)
    out2 = tw.apply_weights_batch(kl, ws, axes=(0, 1, 2))
    assert np.array_equal(np.transpose(out2, (0, 3, 2, 1)), out)                             # layout never changes a value
    with pytest.raises(ValueError):
        tw.apply_weights_batch(k[:, :, :, :-1].copy(), ws, axes=(2, 1, 0))


@pytest.mark.gpu
def test_gpu_sense_combine(gg):
    from mri_acl_imagesegmentation_adsp_b200.fastmri.sense import sens_combine
    g, _ = gg
    img, sens = synth.gaussian_kspace((3, 5, 12, 10), 611), synth.gaussian_kspace((3, 5, 12, 10), 612)
    a = sens_combine(img, sens, magnitude=True)
    assert a.dtype == np.float32 and O.rel_l2(a, g["sense/abs_sum"]) <= 1e-5
    c = sens_combine(img, sens)
    assert c.dtype == np.complex64 and O.rel_l2(c, O.sense_combine(img, sens, False)) <= 1e-5
    shared = sens_combine(torch.from_numpy(img).cuda(), torch.from_numpy(sens[0]).cuda())
    assert O.rel_l2(shared.cpu().numpy(), O.sense_combine(img, sens[:1], False)) <= 1e-5
    with pytest.raises(ValueError):
        sens_combine(img, sens[:, :4])


def test_oracle_t2_reconstruction_vs_golden(gg):
    g, man = gg
    k, calib, hdr = synth.t2_recon_case_inputs()
    from mri_acl_imagesegmentation_adsp_b200.prostate.t2 import _padding_pair, get_padding
    assert get_padding(hdr) == float(g["t2_recon/padding"][0]) and _padding_pair(get_padding(hdr)) == (130, 130)
    rec = O.t2_reconstruction(k, calib, (130, 130))
    assert rec.shape == tuple(man["t2_recon"]["out"]) and O.rel_l2(rec[:, ::2, ::2], g["t2_recon/reconstruction_rss_sub2"]) <= 1e-6


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")
def test_numpy_unravel_bug_makes_the_live_class_drop_holes():
    """numpy 2.3.x: np.unravel_index on an (m, 1)-shaped index array (what np.argwhere returns, grappa.py:88-90) is wrong
    beyond 8192 entries, so the vendored class loses every later hole of a large geometry.  The twin indexes with 1-D
    arrays: it finds all holes, and agrees with the vendored class on every hole the class does find."""
    flat = np.argwhere(np.ones(20000, bool))
    good = np.unravel_index(flat.ravel(), (100, 200))
    bad = np.unravel_index(flat, (100, 200))
    if np.array_equal(bad[0].ravel(), good[0]):
        pytest.skip("this numpy unravels (m, 1) index arrays correctly")
    RG = ref_shim.prostate().grappa.Grappa
    k = synth.gaussian_kspace((150, 2, 160), 5)
    keep = np.zeros(150, bool); keep[::2] = True; keep[70:80] = True
    k[~keep] = 0
    a, b = RG(k.copy(), (5, 5), 1), Grappa(k.copy(), (5, 5), 1)
    true_holes = int((np.abs(k[:, 0, :]) == 0).sum())
    def covered(kv):
        m = np.zeros((150, 160), bool)
        for i in kv["patch_indices"]:
            m[kv["holes_x"][i] - 2, kv["holes_y"][i] - 2] = True
        return m
    ca, cb = covered(a.kernel_var_dict), covered(b.kernel_var_dict)
    assert int(cb.sum()) == true_holes and int(ca.sum()) < true_holes and not (ca & ~cb).any()


@pytest.mark.gpu
def test_gpu_t2_reconstruction_vs_golden(gg):
    """the whole prostate T2 chain (GRAPPA fill of three averages -> pad -> iFFT -> RSS -> flipud -> mean -> crop) as the twin
    of t2_reconstruction, against the vendored function's output on the same inputs."""
    from mri_acl_imagesegmentation_adsp_b200.prostate.t2 import t2_reconstruction
    g, _ = gg
    k, calib, hdr = synth.t2_recon_case_inputs()
    out = t2_reconstruction(k, calib, hdr)
    rec = out["reconstruction_rss"]
    assert set(out) == {"reconstruction_rss"} and rec.shape == (2, 320, 320) and rec.dtype == np.float64
    assert O.rel_l2(rec[:, ::2, ::2], g["t2_recon/reconstruction_rss_sub2"]) <= 1e-5
    dev = t2_reconstruction(torch.from_numpy(k).cuda(), calib, (130, 130))["reconstruction_rss"]
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), rec)
    with pytest.raises(ValueError):
        t2_reconstruction(k[:2], calib, hdr)
