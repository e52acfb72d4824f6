"""Data formats either side of the path: the fastMRI .h5 adapter (through an HDF5-free stand-in, h5py is
not installed) and the per-volume artefact writer, against the reference's own adapter / save_pack where /root/reference exists."""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.adapters.fastmri_adapter import FastMRISinglecoilAdapter, NpzVolumeFile
from mri_acl_imagesegmentation_adsp_b200.dataio.volume_writer import group_records_by_file, preprocess_volumes, save_pack
from oracle import ref_shim


def _make_files(root):
    vols = {}
    for name, n, seed in (("file_b", 4, 1), ("file_a", 3, 2)):
        k = synth.gaussian_kspace((n, 32, 24), seed)
        t = np.abs(k).astype(np.float32)
        extra = {"reconstruction_esc": t} if name == "file_a" else {"reconstruction_rss": t, "reconstruction_esc": t * 2,
                                                                   "mask": (np.arange(24) % 2 == 0).astype(np.float32)}
        NpzVolumeFile.write(os.path.join(root, name + ".h5"), kspace=k, **extra)
        vols[name] = (k, t)
    return vols


def test_adapter_contract(tmp_path):
    vols = _make_files(str(tmp_path))
    ad = FastMRISinglecoilAdapter(str(tmp_path), opener=NpzVolumeFile)
    recs = ad.discover_records()
    assert [(os.path.basename(r["filepath"]), r["slice_idx"]) for r in recs] == \
        [("file_a.h5", i) for i in range(3)] + [("file_b.h5", i) for i in range(4)]          # sorted files, slices in order
    r = ad.load_record(recs[4])
    assert set(r) == {"image", "mask", "label", "kspace", "target", "meta"} and r["image"] is None and r["mask"] is None
    assert np.array_equal(r["kspace"], vols["file_b"][0][1]) and np.array_equal(r["target"], vols["file_b"][1][1])
    assert r["meta"] == {"filepath": recs[4]["filepath"], "slice_idx": 1, "dataset": "fastmri", "target_key": "reconstruction_rss",
                         "adapter": "fastmri_singlecoil-h5"}
    assert ad.load_record(recs[0])["meta"]["target_key"] == "reconstruction_esc"
    assert ad.load_record(recs[4], with_sampling_mask=True)["sampling_mask"].shape == (24,)
    assert ad.load_record(recs[0], with_sampling_mask=True)["sampling_mask"] is None
    assert ad.load_volume_kspace(recs[0]["filepath"]).shape == (3, 32, 24)
    g = group_records_by_file(list(reversed(recs)))
    assert [x["slice_idx"] for x in g[recs[3]["filepath"]]] == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        FastMRISinglecoilAdapter(None, env_key="MRIACL_NO_SUCH_ENV")
    with pytest.raises(ImportError):                                   # default opener: h5py, loudly absent here
        FastMRISinglecoilAdapter(str(tmp_path)).discover_records()


def _pack(seed=0, S=5, H=16, W=12):
    rng = np.random.default_rng(seed)
    mask = (rng.uniform(size=(S, H, W)) > 0.4).astype(np.uint8)
    mask[2] = 0                                                        # an empty mask: NaN statistics for that slice
    return {"tensor": torch.from_numpy(rng.standard_normal((S, 1, H, W)).astype(np.float32)), "preview": rng.uniform(size=(S, H, W)).astype(np.float32),
            "mask": mask, "indices": [7, 8, 9, 10, 11], "sources": ["kspace"] * S, "metas": [{"slice_idx": 7 + i, "filepath": "x.h5"} for i in range(S)]}


def _read_artefacts(d):
    from PIL import Image
    z = np.load(os.path.join(d, "volume.npz"))
    return {"tensor": torch.load(os.path.join(d, "tensor.pt")), "img": z["img"], "msk": z["msk"], "mask": np.load(os.path.join(d, "mask.npy")),
            "indices": json.load(open(os.path.join(d, "indices.json"))), "metas": json.load(open(os.path.join(d, "metas.json"))),
            "stats": json.load(open(os.path.join(d, "stats.json"))),
            "png": {f: np.asarray(Image.open(os.path.join(d, "preview", f))) for f in sorted(os.listdir(os.path.join(d, "preview")))}}


def test_save_pack_artefacts(tmp_path):
    pack = _pack()
    save_pack(str(tmp_path / "vol"), pack, preview_max=3)
    a = _read_artefacts(str(tmp_path / "vol"))
    assert torch.equal(a["tensor"], pack["tensor"]) and a["img"].dtype == np.float32 and a["img"].shape == (5, 1, 16, 12)
    assert a["msk"].dtype == np.uint8 and np.array_equal(a["msk"], pack["mask"]) and np.array_equal(a["mask"], pack["mask"])
    assert a["indices"] == pack["indices"] and a["metas"] == pack["metas"]
    assert sorted(a["png"]) == ["slice_007.png", "slice_008.png", "slice_009.png"]
    assert np.array_equal(a["png"]["slice_008.png"], (pack["preview"][1] * 255).astype(np.uint8))
    st = a["stats"]
    assert st["count_slices"] == 5 and np.isnan(st["per_slice_mean"][2])
    v = pack["tensor"][0, 0].numpy()[pack["mask"][0] > 0]
    assert st["per_slice_mean"][0] == float(v.mean()) and st["per_slice_std"][0] == float(v.std())


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")
def test_formats_vs_live_reference(tmp_path):
    """the reference's own adapter (with the same HDF5 stand-in as its `h5py`) and save_pack on the same inputs."""
    ref_shim.knee_preprocessor_cls()                                   # stubs scikit-image
    from PIL import Image
    for name in ("imageio", "imageio.v2", "tqdm"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["tqdm"].tqdm = lambda x, **k: x
    sys.modules["imageio"].v2 = sys.modules["imageio.v2"]
    sys.modules["imageio.v2"].imwrite = lambda path, arr: Image.fromarray(arr).save(path)
    h5 = sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    h5.File = NpzVolumeFile
    if ref_shim.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_shim.REF_ROOT)
    import importlib
    ref_main = importlib.import_module("src.main")
    ref_adapter = importlib.import_module("src.adapters.fastmri_adapter").FastMRISinglecoilAdapter
    root = tmp_path / "data"
    root.mkdir()
    _make_files(str(root))
    ours, theirs = FastMRISinglecoilAdapter(str(root), opener=NpzVolumeFile), ref_adapter(str(root))
    ra, rb = ours.discover_records(), theirs.discover_records()
    assert ra == rb
    for x, y in zip(ra, rb):
        a, b = ours.load_record(x), theirs.load_record(y)
        assert a["meta"] == b["meta"] and np.array_equal(a["kspace"], b["kspace"]) and np.array_equal(a["target"], b["target"])
    assert group_records_by_file(ra) == ref_main.group_records_by_file(rb)
    pack = _pack(3)
    save_pack(str(tmp_path / "ours"), pack, preview_max=4)
    ref_main.save_pack(str(tmp_path / "theirs"), pack, preview_max=4)
    a, b = _read_artefacts(str(tmp_path / "ours")), _read_artefacts(str(tmp_path / "theirs"))
    assert torch.equal(a["tensor"], b["tensor"]) and np.array_equal(a["img"], b["img"]) and np.array_equal(a["msk"], b["msk"])
    assert np.array_equal(a["mask"], b["mask"]) and a["indices"] == b["indices"] and a["metas"] == b["metas"]
    assert sorted(a["png"]) == sorted(b["png"]) and all(np.array_equal(a["png"][k], b["png"][k]) for k in a["png"])
    assert json.dumps(a["stats"], sort_keys=True) == json.dumps(b["stats"], sort_keys=True)
    for f in ("tensor.pt", "mask.npy", "indices.json", "metas.json", "stats.json"):           # byte-identical files
        assert open(tmp_path / "ours" / f, "rb").read() == open(tmp_path / "theirs" / f, "rb").read(), f


@pytest.mark.gpu
def test_files_to_artefacts_on_the_device(tmp_path):
    """.h5 stand-ins -> adapter -> device preprocessor (recon + clip + resize + z-score) -> artefacts the trainer reads."""
    from mri_acl_imagesegmentation_adsp_b200.preprocess.mri_preprocess import MRIKneePreprocessor
    from oracle import recon_oracle as O
    root = tmp_path / "data"
    root.mkdir()
    vols = _make_files(str(root))
    pre = MRIKneePreprocessor(out_size=(16, 16), slice_keep=(0.0, 1.0))
    summary = preprocess_volumes(FastMRISinglecoilAdapter(str(root), opener=NpzVolumeFile), str(tmp_path / "out"), pre)
    assert [s["num_slices"] for s in summary] == [3, 4]
    a = _read_artefacts(summary[1]["output_dir"])
    assert a["img"].shape == (4, 1, 16, 16) and a["indices"] == [0, 1, 2, 3]
    assert json.load(open(os.path.join(summary[1]["output_dir"], "metas.json")))[0]["target_key"] == "reconstruction_rss"
    t = vols["file_b"][1]                      # the files carry a reconstruction: it wins over k-space (mri_preprocess.py:262-276)
    for s in range(4):
        z, p01, mr, _ = O.post_chain(t[s], None, (16, 16), (1.0, 99.5))
        assert np.abs(a["img"][s, 0] - z).max() <= 2e-6 * max(1.0, float(np.abs(z).max())) and np.array_equal(a["msk"][s], mr)
    # k-space only records (no reconstruction in the file): the device reconstruction feeds the same steps
    k = vols["file_b"][0]
    recs = [{"kspace": k[s], "meta": {"slice_idx": s}} for s in range(4)]
    out = pre.preprocess_records(recs)
    assert out["sources"] == ["kspace"] * 4
    for s in range(4):
        z, _, _, _ = O.post_chain(O.ifft2c_single(k[s]), None, (16, 16), (1.0, 99.5))
        assert np.abs(out["tensor"][s, 0].numpy() - z).max() <= 2e-6 * max(1.0, float(np.abs(z).max()))
