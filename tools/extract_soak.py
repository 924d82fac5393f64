"""Randomised soak of the extractor against the oracle: odd image sizes, scale factors, level counts, thresholds, masks, strided inputs.
usage: python tools/extract_soak.py [n_cases]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import orbslam_mapsave_b200 as orb
from orbslam_mapsave_b200.synth import synth
from oracle import orb_oracle_py as orc

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
bad = 0
tot_kp = 0
for case in range(n_cases):
    rng = np.random.default_rng(20_000 + case)
    sf = float(rng.choice([1.1, 1.2, 1.2, 1.2, 1.3, 1.5, 2.0]))
    nl = int(rng.integers(1, 9))
    while 70 * sf ** (nl - 1) > 500:                   # every level must stay >= 62 px (reference precondition)
        nl -= 1
    smallest = int(70 * sf ** (nl - 1)) + 10
    W = int(rng.integers(smallest, 1100))
    H = int(rng.integers(max(smallest, W // 2 + 2), max(smallest, W // 2 + 2, min(W * 2 - 2, 900)) + 1))
    nf = int(rng.choice([100, 500, 1000, 2000, 4000]))
    ini, mn = (20, 7) if rng.random() < 0.7 else (int(rng.integers(8, 60)), int(rng.integers(3, 8)))
    img = synth(W, H, 30_000 + case)
    mask = None
    if rng.random() < 0.3:
        mask = np.full((H, W), 255, np.uint8)
        x0, y0 = int(rng.integers(0, W // 2)), int(rng.integers(0, H // 2))
        mask[y0:y0 + H // 3, x0:x0 + W // 3] = 0
    try:
        ex = orb.ORBextractor(nf, sf, nl, ini, mn)
        kp, d = ex(img, mask)
        oex = orc.Extractor(nf, sf, nl, ini, mn)
        okp, od = oex.extract(img, mask)
        same = len(kp) == len(okp) and all(np.array_equal(kp[f], okp[f]) for f in ("x", "y", "octave", "response", "size"))
        pyr_ok = all(np.array_equal(ex.pyramid_level(0, l), oex.level(l)) for l in range(nl))
        ang_ok = same and (len(kp) == 0 or np.abs(kp["angle"] - okp["angle"]).max() <= 1e-3)
        bits = int(np.unpackbits(d ^ od).sum()) if same and len(kp) else 0
        if not (same and pyr_ok and ang_ok and bits == 0):
            bad += 1
            print("MISMATCH case", case, (W, H, nf, nl, sf, ini, mn), "kp", same, "pyr", pyr_ok, "ang", ang_ok, "bits", bits, flush=True)
        tot_kp += len(okp)
    except orb.OrbError as e:
        print("case", case, (W, H, nf, nl, sf), "->", str(e)[:80], flush=True)
print(f"extract soak: {n_cases} cases, {tot_kp} keypoints, {bad} mismatching", flush=True)
