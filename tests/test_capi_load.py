"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/orb_b200.h declares,
and compute entry points fail loudly (no CPU fallback) when there is no CUDA device."""
import os
import re

import numpy as np
import pytest

import orbslam_mapsave_b200 as orb
from orbslam_mapsave_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "orb_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(orb(?:[xmv]|map)?_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    L = orb.lib()
    for s in declared:
        assert getattr(L, s) is not None


def test_keypoint_struct_is_28_bytes():
    assert capi.KP_DTYPE.itemsize == 28


@pytest.mark.skipif(orb.device_count() > 0, reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(orb.OrbError):
        orb.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480)
    with pytest.raises(orb.OrbError):
        orb.ORBmatcher().hamming_top2(np.zeros((4, 32), np.uint8), np.zeros((4, 32), np.uint8))
    with pytest.raises(orb.OrbError):
        orb.ORBmatcher.DescriptorDistance(np.zeros(32, np.uint8), np.zeros(32, np.uint8))
    # the widened rows fail the same way: window searches, vocabulary
    n = 8
    g = orb.GridView(np.zeros((n, 32), np.uint8), np.linspace(10, 600, n), np.linspace(10, 400, n), np.zeros(n, np.int32),
                     np.ones(8, np.float32), (0.0, 0.0, 640.0, 480.0), angle=np.zeros(n, np.float32))
    assert g.cell_offsets[-1] == n and len(g.cell_offsets) == 64 * 48 + 1            # host-side grid build needs no GPU
    with pytest.raises(orb.OrbError):
        orb.ORBmatcher().SearchForInitialization(g, np.zeros((n, 32), np.uint8), np.zeros(n, np.int32), np.zeros(n, np.float32),
                                                 np.zeros((n, 2), np.float32), 100)
    with pytest.raises(orb.OrbError):
        orb.ORBmatcher().SearchWindowsBest(g, np.ones(n, np.uint8), np.zeros(n), np.zeros(n), np.ones(n), np.zeros(n, np.int32),
                                           np.zeros(n, np.int32), np.zeros((n, 32), np.uint8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "orbslam_mapsave_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "orb_oracle" not in txt and "liborb_oracle" not in txt, f


def test_constructor_tables_need_no_device():
    """orbx_compute_tables = the reference constructor's tables (src/ORBextractor.cc:414-445) without a handle or a CUDA device: equal
    to the oracle's (which equal the reference build's, tests/test_reference_ref.py)."""
    import numpy as np
    from oracle import orb_oracle_py as orc
    from orbslam_mapsave_b200 import capi
    for nf, sf, nl in [(1000, 1.2, 8), (2000, 1.2, 8), (8000, 1.2, 12), (500, 1.5, 3), (1200, 2.0, 5), (777, 1.1, 16), (10, 1.2, 1)]:
        t = [np.zeros(nl, np.float32) for _ in range(4)] + [np.zeros(nl, np.int32)]
        capi.check(capi.lib().orbx_compute_tables(nf, sf, nl, *[capi._p(a) for a in t]))
        o = orc.Extractor(nf, sf, nl, 20, 7).tables()
        for got, key in zip(t, ("scale", "inv_scale", "sigma2", "inv_sigma2", "quota")):
            assert np.array_equal(got, o[key]), (nf, sf, nl, key)
    import orbslam_mapsave_b200 as orb
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7)                  # getters before any image: no device is needed
    assert ex.GetScaleFactors()[1] == np.float32(1.2) and ex.features_per_level().sum() == 1000


def test_host_sources_compile_against_an_opencv_shaped_header(tmp_path):
    """The drop-in classes compiled the way they are inside the reference's tree: cv_compat.h steps aside (`__has_include`) for an
    <opencv2/core/core.hpp> — here tests/cpp/opencv_layout, a header with the real cv::Mat member order and the _InputArray /
    _OutputArray proxies (OpenCV itself is not in this image) — so the sources use nothing that only the stand-in offers."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "orbslam_mapsave_b200", "host")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    for src in ("ORBextractor.cc", "ORBmatcher.cc", "ORBVocabulary.cc", "MapArchive.cc"):
        r = subprocess.run([cxx, "-O0", "-std=c++14", "-fPIC", "-Wall", "-Werror", "-ffp-contract=off", "-DORB_B200_FORCE_MIN_TYPES",
                            "-I" + os.path.join(root, "tests", "cpp", "opencv_layout"), "-c", "-o", str(tmp_path / (src + ".o")),
                            os.path.join(host, src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
