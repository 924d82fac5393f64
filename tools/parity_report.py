"""Parity statistics the north_star asks to be REPORTED: over many synthetic frames, GPU (through the C ABI) vs the CPU oracle:
keypoint (x, y, octave, response) set equality, max angle difference, descriptor bit mismatch rate.
usage: python tools/parity_report.py [n_frames_640x480] > profiles/<name>.json"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import orbslam_mapsave_b200 as orb
from orbslam_mapsave_b200.synth import synth
from oracle import orb_oracle_py as orc


def run(W, H, nf, nl, seeds):
    ex = orb.ORBextractor(nf, 1.2, nl, 20, 7)
    oex = orc.Extractor(nf, 1.2, nl, 20, 7)
    tot = dict(frames=0, keypoints=0, frames_with_keypoint_set_mismatch=0, max_abs_angle_diff_deg=0.0, angle_bits_different=0,
               descriptor_bits=0, descriptor_bits_different=0)
    for s in seeds:
        img = synth(W, H, s)
        kp, d = ex(img, download_pyramid=False)
        okp, od = oex.extract(img)
        tot["frames"] += 1
        same = len(kp) == len(okp) and all(np.array_equal(kp[f], okp[f]) for f in ("x", "y", "octave", "response", "size"))
        if not same:
            tot["frames_with_keypoint_set_mismatch"] += 1
            continue
        tot["keypoints"] += len(kp)
        da = np.abs(kp["angle"].astype(np.float64) - okp["angle"].astype(np.float64))
        da = np.minimum(da, 360.0 - da)
        tot["max_abs_angle_diff_deg"] = max(tot["max_abs_angle_diff_deg"], float(da.max()) if len(da) else 0.0)
        tot["angle_bits_different"] += int((kp["angle"].view(np.uint32) != okp["angle"].view(np.uint32)).sum())
        tot["descriptor_bits"] += d.size * 8
        tot["descriptor_bits_different"] += int(np.unpackbits(d ^ od).sum())
    tot["descriptor_bit_mismatch_rate"] = tot["descriptor_bits_different"] / max(tot["descriptor_bits"], 1)
    return tot


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    out = {"C1/C2 640x480 nF=1000 8 levels": run(640, 480, 1000, 8, range(n)),
           "C3 1280x720 nF=2000 8 levels": run(1280, 720, 2000, 8, range(5000, 5000 + max(n // 8, 4))),
           "C5 3840x2160 nF=8000 12 levels": run(3840, 2160, 8000, 12, range(9000, 9003)),
           "bars": {"keypoint sets": "bit-exact", "angle": "<= 1e-3 deg", "descriptor bits": "<= 1e-4 mismatching"}}
    print(json.dumps(out, indent=1))
