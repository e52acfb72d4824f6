"""The built shared library loads (no GPU needed) and exports every symbol the header declares."""
import os
import re
import subprocess

import pytest

from mri_acl_imagesegmentation_adsp_b200.adapters import recon_cabi as cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mriacl_recon.h")


@pytest.fixture(scope="module")
def built():
    if not os.path.isfile(cabi.DEFAULT_LIBRARY):
        subprocess.run(["make", "-s", "-C", os.path.dirname(cabi.DEFAULT_LIBRARY)], check=True)
    return cabi.ReconLibrary(cabi.DEFAULT_LIBRARY)


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mriacl_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(built):
    names = header_functions()
    assert names, "no declarations found"
    assert sorted(cabi.SIGNATURES) == names
    for n in names:
        assert hasattr(built._lib, n)


def test_queries_without_gpu(built):
    assert built._lib.mriacl_abi_version() == cabi.ABI_VERSION
    assert built.supported(640, 368) == cabi.PATH_FUSED
    assert built.supported(640, 640) == cabi.PATH_FUSED        # prostate-shape plan (rowpass640)
    assert built.supported(640, 372) == cabi.PATH_FUSED        # 372 = 31 x 12 knee plan
    assert built.supported(640, 320) == cabi.PATH_GENERIC      # any-width path (pruned behind the 640 column pass)
    assert built.supported(372, 640) == cabi.PATH_GENERIC
    assert built.supported(640, 5000) == cabi.PATH_NONE
    n1 = built.recon_rss_workspace_bytes(1, 1, 15, 640, 368, 0, 368, 320, 320)
    n4 = built.recon_rss_workspace_bytes(4, 1, 15, 640, 368, 0, 368, 320, 320)
    assert n4 == 4 * n1 and n1 >= 15 * 368 * 320 * 8
    with pytest.raises(ValueError):
        built.recon_rss_workspace_bytes(1, 1, 15, 640, 368, 0, 368, 641, 320)
    assert built.launch_count() == 0


def test_missing_library_is_loud(tmp_path):
    with pytest.raises(cabi.ReconLibraryError):
        cabi.ReconLibrary(str(tmp_path / "nope.so"))


def test_is_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "-lelf", built.path], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
