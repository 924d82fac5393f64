"""GPU parity tests of the extractor path (through the C ABI) against the CPU oracle and the cv2 golden fixtures.

Bars (BASELINE.json north_star): keypoint (x, y, octave, response) bit-exact; angles within 1e-3 deg; descriptor bit
mismatch rate <= 0.01 %.  Intermediate stages (pyramid, FAST candidates, blur) are integer work and must be bit-exact.
"""
import glob
import os

import numpy as np
import pytest

import orbslam_mapsave_b200 as orb
from orbslam_mapsave_b200.synth import synth
from oracle import orb_oracle_py as orc

pytestmark = pytest.mark.gpu
ANGLE_TOL_DEG = 1e-3
DESC_BIT_MISMATCH_MAX = 1e-4      # 0.01 %


def _compare(gpu_kp, gpu_desc, ref_kp, ref_desc):
    assert len(gpu_kp) == len(ref_kp)
    for f in ("x", "y", "octave", "response", "size", "class_id"):
        assert np.array_equal(gpu_kp[f], ref_kp[f]), f
    dang = np.abs(gpu_kp["angle"].astype(np.float64) - ref_kp["angle"].astype(np.float64))
    dang = np.minimum(dang, 360.0 - dang)
    assert dang.max(initial=0.0) <= ANGLE_TOL_DEG
    bits = np.unpackbits(gpu_desc ^ ref_desc).sum()
    rate = bits / max(1, gpu_desc.size * 8)
    assert rate <= DESC_BIT_MISMATCH_MAX, rate
    return int((gpu_kp["angle"] != ref_kp["angle"]).sum()), int(bits)


def _stages_equal(ex, oex, nlevels):
    for l in range(nlevels):
        assert np.array_equal(ex.pyramid_level(0, l), oex.level(l)), f"pyramid level {l}"
        assert np.array_equal(ex.pyramid_level(0, l, bordered=True), oex.level(l, bordered=True)), f"bordered level {l}"
        got, want = ex.candidates(0, l), oex.candidates(l)
        assert len(got) == len(want), (l, len(got), len(want))
        for f in ("x", "y", "r"):
            assert np.array_equal(got[f], want[f]), (l, f)
        b = oex.blurred(l)
        if b is not None:
            assert np.array_equal(ex.blurred_level(0, l), b), f"blurred level {l}"


def _chain_image(g, W, H, seed):
    """The fixture's image, or (large cases) the synthetic frame regenerated from its seed and checked against the recorded CRC."""
    if "image" in g.files:
        return g["image"]
    import zlib
    from orbslam_mapsave_b200.synth import synth as _synth
    img = _synth(W, H, seed)
    assert zlib.crc32(img.tobytes()) == int(g["image_crc"]), "synth() no longer reproduces the frame this fixture was made from"
    return img


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "chain_*.npz"))),
                         ids=lambda p: os.path.basename(p)[6:-4])
def test_against_cv2_chain_golden(path):
    g = np.load(path)
    W, H, seed, nf, nl, ini, mn, use_mask = (int(v) for v in g["params"])
    img = _chain_image(g, W, H, seed)
    mask = None
    if use_mask:
        mask = np.full(img.shape, 255, np.uint8)
        mask[60:220, 150:260] = 0
    ex = orb.ORBextractor(nf, float(g["scaleFactor"]), nl, ini, mn)
    kp, desc = ex(img, mask)
    assert np.array_equal(ex.features_per_level(), g["quota"])
    assert np.array_equal(ex.pyramid_level(0, nl - 1), g["last_level"])
    assert [len(ex.candidates(0, l)) for l in range(nl)] == g["cand_counts"].tolist()
    _compare(kp, desc, g["kp"], g["desc"])


@pytest.mark.parametrize("W,H,seed,nf,nl,sf,ini,mn", [
    (640, 480, 1, 1000, 8, 1.2, 20, 7),          # C1/C2 geometry
    (640, 480, 2, 2000, 8, 1.2, 32, 7),          # the fork's shipped ORB_RGB640x480.yaml values
    (1280, 720, 3, 2000, 8, 1.2, 20, 7),         # C3 (nIni = 2 root nodes)
    (424, 240, 4, 300, 3, 1.5, 15, 3),           # ORB_RGBD640x480.yaml-style scale 1.5
    (200, 150, 5, 100, 4, 1.2, 20, 7),           # tiny: single-column cell grids at the top levels
    (487, 641, 6, 400, 5, 1.3, 25, 5),           # portrait, odd sizes
])
def test_stage_by_stage_vs_oracle(W, H, seed, nf, nl, sf, ini, mn):
    img = synth(W, H, seed)
    ex = orb.ORBextractor(nf, sf, nl, ini, mn)
    kp, desc = ex(img)
    oex = orc.Extractor(nf, sf, nl, ini, mn)
    okp, odesc = oex.extract(img)
    t = oex.tables()
    assert np.array_equal(ex.GetScaleFactors(), t["scale"])
    assert np.array_equal(ex.GetInverseScaleFactors(), t["inv_scale"])
    assert np.array_equal(ex.GetScaleSigmaSquares(), t["sigma2"])
    assert np.array_equal(ex.GetInverseScaleSigmaSquares(), t["inv_sigma2"])
    assert np.array_equal(ex.features_per_level(), t["quota"])
    _stages_equal(ex, oex, nl)
    _compare(kp, desc, okp, odesc)


def test_edge_inputs():
    ex = orb.ORBextractor(500, 1.2, 4, 20, 7)
    # flat image: no corners anywhere -> zero keypoints, descriptors released
    kp, desc = ex(np.full((240, 320), 77, np.uint8))
    assert len(kp) == 0 and desc.shape == (0, 32)
    # pure noise: far more candidates than the quota on every level (octree cull + dense cells)
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (240, 320), dtype=np.uint8)
    kp, desc = ex(img)
    okp, odesc = orc.Extractor(500, 1.2, 4, 20, 7).extract(img)
    _compare(kp, desc, okp, odesc)
    # fully masked-out image == black image
    kp, desc = ex(img, np.zeros_like(img))
    assert len(kp) == 0
    # empty image: silent no-op like the reference (ORBextractor.cc:1045-1046)
    assert ex(np.zeros((0, 0), np.uint8)) is None
    # a level below 62 px is a reference precondition violation -> explicit error instead of a division by zero
    with pytest.raises(orb.OrbError):
        orb.ORBextractor(100, 1.2, 8, 20, 7)(np.zeros((100, 100), np.uint8))
    # aspect ratio < 0.5 gives nIni = round(w/h) = 0 root nodes: the reference divides by zero (ORBextractor.cc:542-544)
    with pytest.raises(orb.OrbError):
        orb.ORBextractor(100, 1.2, 2, 20, 7)(np.zeros((777, 333), np.uint8))


def test_strided_input_and_mask():
    big = synth(700, 500, 11)
    view = big[10:490, 30:670]                         # non-contiguous rows
    mask = np.full((480, 640), 255, np.uint8)
    mask[100:300, 200:400] = 0
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7)
    kp, desc = ex(view, mask)
    okp, odesc = orc.Extractor(1000, 1.2, 8, 20, 7).extract(view, mask)
    _compare(kp, desc, okp, odesc)
    inside = (kp["octave"] == 0) & (kp["x"] > 210) & (kp["x"] < 390) & (kp["y"] > 110) & (kp["y"] < 290)
    assert not inside.any()


def test_batch_matches_single_and_oracle():
    frames = np.stack([synth(640, 480, s) for s in range(20, 26)])
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=4)          # 6 frames through a 4-frame workspace: 2 passes
    kp, desc, n = ex.extract_batch(frames)
    oex = orc.Extractor(1000, 1.2, 8, 20, 7)
    nang = nbits = 0
    for f in range(len(frames)):
        okp, odesc = oex.extract(frames[f])
        a, b = _compare(kp[f, :n[f]], desc[f, :n[f]], okp, odesc)
        nang += a
        nbits += b
    print(f"angle float mismatches: {nang}, descriptor bit mismatches: {nbits} over {int(n.sum())} keypoints")


def test_ramped_chunk_schedule_matches_single_calls():
    """orbx_extract_batch over many passes uses small first / last chunks (64, 128, 128, 16, 64 frames here) to shorten the
    pipeline's fill and drain, over four staging slots; every frame must come back in its own slot, identical to a single-frame call."""
    base = [synth(640, 480, s) for s in range(40, 50)]
    frames = np.stack([np.roll(base[i % 10], 3 * (i // 10), axis=1) for i in range(400)])
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=128)
    kp, desc, n = ex.extract_batch(frames)
    one = orb.ORBextractor(1000, 1.2, 8, 20, 7)
    for f in range(len(frames)):
        k1, d1 = one(frames[f])
        assert n[f] == len(k1), f
        assert np.array_equal(kp[f, :n[f]].view(np.uint32), k1.view(np.uint32).reshape(-1)) or \
            np.array_equal(kp[f, :n[f]].tobytes(), k1.tobytes()), f
        assert np.array_equal(desc[f, :n[f]], d1), f
    okp, odesc = orc.Extractor(1000, 1.2, 8, 20, 7).extract(frames[137])
    _compare(kp[137, :n[137]], desc[137, :n[137]], okp, odesc)


def test_pageable_strided_and_pinned_pipelines_agree():
    """orbx_extract_batch with slot reuse (700 frames in 64-frame chunks over six slots): ordinary numpy memory (host-pool staging in and
    out), a strided view of a wider array (rows packed by the pool) and page-locked buffers (direct DMA) must give identical results."""
    import ctypes as C
    import torch
    from orbslam_mapsave_b200 import capi
    base = [synth(640, 480, s) for s in range(60, 67)]
    frames = np.stack([np.roll(base[i % 7], 5 * (i // 7), axis=0) for i in range(700)])
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=64)
    kp, desc, n = ex.extract_batch(frames)                               # pageable in, pageable out
    cap = ex.max_keypoints()
    wide = np.zeros((700, 483, 656), np.uint8)
    wide[:, :480, :640] = frames
    kp2 = np.zeros((700, cap), capi.KP_DTYPE)
    desc2 = np.zeros((700, cap, 32), np.uint8)
    n2 = np.zeros(700, np.int32)
    capi.check(capi.lib().orbx_extract_batch(ex.handle, capi._p(wide), 700, 640, 480, 656, 483 * 656, None, 0, 0, capi._p(kp2), capi._p(desc2),
                                             cap, capi._p(n2)))
    h_frames = torch.from_numpy(frames).pin_memory()
    h_kp = torch.zeros((700, cap, 7), dtype=torch.float32).pin_memory()
    h_desc = torch.zeros((700, cap, 32), dtype=torch.uint8).pin_memory()
    n3 = np.zeros(700, np.int32)
    capi.check(capi.lib().orbx_extract_batch(ex.handle, capi._p(h_frames), 700, 640, 480, 640, 640 * 480, None, 0, 0, capi._p(h_kp),
                                             capi._p(h_desc), cap, capi._p(n3)))
    kp3 = h_kp.numpy().view(np.uint32)
    assert np.array_equal(n, n2) and np.array_equal(n, n3) and n.min() > 900
    for f in range(700):
        a = kp[f, :n[f]].view(np.uint32).reshape(-1, 7)
        assert np.array_equal(a, kp2[f, :n[f]].view(np.uint32).reshape(-1, 7)), f
        assert np.array_equal(a, kp3[f, :n[f]]), f
        assert np.array_equal(desc[f, :n[f]], desc2[f, :n[f]]) and np.array_equal(desc[f, :n[f]], h_desc[f, :n[f]].numpy()), f
    one = orb.ORBextractor(1000, 1.2, 8, 20, 7)
    for f in (0, 63, 64, 389, 699):
        k1, d1 = one(frames[f])
        assert n[f] == len(k1) and np.array_equal(kp[f, :n[f]].tobytes(), k1.tobytes()) and np.array_equal(desc[f, :n[f]], d1), f


def test_batch_of_separately_allocated_frames():
    """orbx_extract_batch_ptrs (frames as separate allocations, like std::vector<cv::Mat>): pageable arrays through the staging pool,
    strided views, page-locked tensors by per-frame DMA, and the single-pass (latency) form all equal orbx_extract_batch."""
    import torch
    base = [synth(640, 480, s) for s in range(80, 85)]
    frames = [np.roll(base[i % 5], 7 * (i // 5), axis=1).copy() for i in range(300)]
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=64)
    kp0, desc0, n0 = ex.extract_batch(np.stack(frames))
    kp1, desc1, n1 = ex.extract_batch_list(frames)                       # pageable, 5 chunks
    wide = [np.zeros((480, 656), np.uint8) for _ in range(300)]
    for w_, f_ in zip(wide, frames):
        w_[:, :640] = f_
    kp2, desc2, n2 = ex.extract_batch_list([w_[:, :640] for w_ in wide])   # row stride 656
    pinned = [torch.from_numpy(f_).pin_memory() for f_ in frames[:100]]
    kp3, desc3, n3 = ex.extract_batch_list([t.numpy() for t in pinned])    # page-locked frames: one 2-D DMA each
    kp4, desc4, n4 = ex.extract_batch_list(frames[:9])                     # one pass (the latency path)
    assert np.array_equal(n0, n1) and np.array_equal(n0, n2) and np.array_equal(n0[:100], n3) and np.array_equal(n0[:9], n4) and n0.min() > 900
    for f in range(300):
        a, d = kp0[f, :n0[f]].tobytes(), desc0[f, :n0[f]]
        assert a == kp1[f, :n0[f]].tobytes() and np.array_equal(d, desc1[f, :n0[f]]), f
        assert a == kp2[f, :n0[f]].tobytes() and np.array_equal(d, desc2[f, :n0[f]]), f
        if f < 100:
            assert a == kp3[f, :n0[f]].tobytes() and np.array_equal(d, desc3[f, :n0[f]]), f
        if f < 9:
            assert a == kp4[f, :n0[f]].tobytes() and np.array_equal(d, desc4[f, :n0[f]]), f


def test_non_tma_fallback_kernels_match_oracle():
    """ORBX_FAST_TMA=0 / ORBX_DESC_TMA=0 select k_fast (one CTA per cell) and k_describe<false> (cp.async staging) — the kernels that
    run when tensor maps cannot be encoded or cells exceed the TMA box.  Separate process: the switches are read when the library plans."""
    import subprocess
    import sys
    code = (
        "import numpy as np, orbslam_mapsave_b200 as orb\n"
        "from orbslam_mapsave_b200.synth import synth\n"
        "from oracle import orb_oracle_py as orc\n"
        "for (W, H, seed, nf, nl) in ((640, 480, 3, 1000, 8), (400, 300, 5, 500, 5)):\n"
        "    img = synth(W, H, seed)\n"
        "    kp, desc = orb.ORBextractor(nf, 1.2, nl, 20, 7)(img)\n"
        "    okp, odesc = orc.Extractor(nf, 1.2, nl, 20, 7).extract(img)\n"
        "    assert len(kp) == len(okp) and len(kp) > 300\n"
        "    for f in ('x', 'y', 'octave', 'response', 'size'):\n"
        "        assert np.array_equal(kp[f], okp[f]), f\n"
        "    assert np.abs(kp['angle'] - okp['angle']).max() <= 1e-3\n"
        "    assert np.unpackbits(desc ^ odesc).sum() <= 1e-4 * desc.size * 8\n"
        "frames = np.stack([synth(640, 480, s) for s in range(8)])\n"
        "ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=4)\n"
        "k, d, n = ex.extract_batch(frames)\n"
        "one = orb.ORBextractor(1000, 1.2, 8, 20, 7)\n"
        "for f in range(8):\n"
        "    k1, d1 = one(frames[f])\n"
        "    assert n[f] == len(k1) and np.array_equal(k[f, :n[f]].tobytes(), k1.tobytes()) and np.array_equal(d[f, :n[f]], d1)\n"
        "print('fallback ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for env in ({"ORBX_FAST_TMA": "0"}, {"ORBX_DESC_TMA": "0"}, {"ORBX_BLUR_FORK": "0", "ORBX_FW_WARPS": "5", "ORBX_DESC_WARPS": "7", "ORBX_DESC_CHUNK": "3"}):
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "fallback ok" in r.stdout, (env, r.stdout[-400:], r.stderr[-1200:])


def test_device_resident_multi_pass_matches_single_calls():
    """orbx_extract_batch_device over more frames than the workspace holds runs several passes; results must land in the right
    rows and equal single-frame calls, and later work on the caller's stream must see them."""
    import torch
    base = [synth(640, 480, s) for s in range(60, 65)]
    host = np.stack([np.roll(base[i % 5], 2 * (i // 5), axis=0) for i in range(22)])
    frames = torch.from_numpy(host).cuda()
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=4)            # 22 frames: 6 passes, the last one ragged
    ex._plan(640, 480)
    cap = ex.max_keypoints()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        kp = torch.zeros((22, cap, 7), dtype=torch.float32, device="cuda")
        desc = torch.zeros((22, cap, 32), dtype=torch.uint8, device="cuda")
        n = torch.zeros(22, dtype=torch.int32, device="cuda")
        ex.extract_batch_device(frames, kp, desc, n, cap, stream=st.cuda_stream)
        n2 = n.clone()                                                 # ordered after the passes on the same stream
    st.synchronize()
    ex.check_status()
    kp, desc, n = kp.cpu().numpy(), desc.cpu().numpy(), n2.cpu().numpy()
    one = orb.ORBextractor(1000, 1.2, 8, 20, 7)
    for f in range(22):
        k1, d1 = one(host[f])
        assert n[f] == len(k1) and n[f] > 900, f
        assert kp[f, :n[f]].tobytes() == k1.tobytes(), f
        assert np.array_equal(desc[f, :n[f]], d1), f


def test_device_resident_batch_idempotent():
    import torch
    frames = torch.from_numpy(np.stack([synth(640, 480, s) for s in range(30, 34)])).cuda()
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=4)
    ex._plan(640, 480)
    cap = ex.max_keypoints()
    outs = []
    for _ in range(2):
        kp = torch.zeros((4, cap, 7), dtype=torch.float32, device="cuda")
        desc = torch.zeros((4, cap, 32), dtype=torch.uint8, device="cuda")
        n = torch.zeros(4, dtype=torch.int32, device="cuda")
        ex.extract_batch_device(frames, kp, desc, n, cap, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ex.check_status()
        outs.append((kp.cpu().numpy().copy(), desc.cpu().numpy().copy(), n.cpu().numpy().copy()))
    assert np.array_equal(outs[0][2], outs[1][2])
    for f in range(4):
        k = outs[0][2][f]
        assert np.array_equal(outs[0][0][f, :k].view(np.uint32), outs[1][0][f, :k].view(np.uint32))
        assert np.array_equal(outs[0][1][f, :k], outs[1][1][f, :k])
    okp, odesc = orc.Extractor(1000, 1.2, 8, 20, 7).extract(frames[2].cpu().numpy())
    k = outs[0][2][2]
    gkp = outs[0][0][2, :k].copy().view(orb.KP_DTYPE).reshape(-1)
    _compare(gkp, outs[0][1][2, :k], okp, odesc)


@pytest.mark.parametrize("W,H,nf,nl", [(3840, 2160, 8000, 12)])
def test_4k_config5(W, H, nf, nl):
    img = synth(W, H, 0)
    ex = orb.ORBextractor(nf, 1.2, nl, 20, 7)
    kp, desc = ex(img, download_pyramid=False)
    oex = orc.Extractor(nf, 1.2, nl, 20, 7)
    okp, odesc = oex.extract(img)
    for l in (0, 5, 11):
        got, want = ex.candidates(0, l), oex.candidates(l)
        assert np.array_equal(got["x"], want["x"]) and np.array_equal(got["y"], want["y"]) and np.array_equal(got["r"], want["r"])
    _compare(kp, desc, okp, odesc)
