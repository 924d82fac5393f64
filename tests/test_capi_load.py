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
