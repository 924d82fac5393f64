"""CPU tests of the map-archive reader / writer (orbmap_* of include/orb_b200.h) against the independent Python
statement of the fork's Boost binary archive layout (tests/map_archive_writer.py).  No GPU needed: this is host-side format
code.  Layout parity with a real Boost build is unpinned (no Boost in this image, no .bin shipped by the reference)."""
import os

import numpy as np
import pytest

import orbslam_mapsave_b200 as orb
from orbslam_mapsave_b200.capi import OrbError
from map_archive_writer import make_random_map, serialize_map


def _ids(v):
    return np.array([-1 if (x is None or x < 0) else x for x in v], np.int64)


def _check_keyframe(got, k):
    inf = got["info"]
    for f, key in (("id", "id"), ("frame_id", "frame_id"), ("next_id", "next_id"), ("n", "n"), ("n_levels", "n_levels"),
                   ("grid_cols", "grid_cols"), ("grid_rows", "grid_rows"), ("min_x", "min_x"), ("max_x", "max_x"), ("min_y", "min_y"),
                   ("max_y", "max_y")):
        assert inf[f] == k[key], f
    assert inf["timestamp"] == k["timestamp"]
    for f in ("fx", "fy", "cx", "cy", "invfx", "invfy", "bf", "b", "th_depth", "scale_factor", "log_scale_factor", "grid_inv_w",
              "grid_inv_h", "half_baseline"):
        assert np.float32(inf[f]) == np.float32(k[f]), f
    assert inf["has_parent"] == (k["parent"] is not None and k["parent"] >= 0)
    if inf["has_parent"]:
        assert inf["parent_id"] == k["parent"]
    assert (inf["is_bad"], inf["not_erase"], inf["to_be_erased"], inf["first_connection"]) == (
        int(k["bad"]), int(k["not_erase"]), int(k["to_be_erased"]), int(k["first_connection"]))
    for name in ("keys", "keys_un"):
        a = got[name]
        assert len(a) == len(k[name])
        for f in ("x", "y", "angle", "response", "octave", "class_id"):
            assert np.array_equal(a[f], np.array([q[f] for q in k[name]], a[f].dtype)), (name, f)
        assert not a["size"].any()          # never stored by the fork
    assert np.array_equal(got["uright"], k["uright"]) and np.array_equal(got["depth"], k["depth"])
    assert np.array_equal(got["desc"], k["desc"])
    assert np.array_equal(got["mappoint_ids"], _ids(k["mappoint_ids"]))
    assert np.array_equal(got["scale_factors"], k["scale_factors"]) and np.array_equal(got["level_sigma2"], k["level_sigma2"])
    assert np.array_equal(got["inv_level_sigma2"], k["inv_level_sigma2"])
    assert np.array_equal(got["Tcw"], k["Tcw"]) and np.array_equal(got["K"], k["K"])
    assert np.array_equal(got["connected_ids"], _ids([c[0] for c in k["connected"]]))
    assert np.array_equal(got["connected_weights"], np.array([0 if c[0] < 0 else c[1] for c in k["connected"]], np.int32))
    assert np.array_equal(got["ordered_ids"], _ids(k["ordered_ids"])) and np.array_equal(got["ordered_weights"], k["ordered_weights"])
    assert np.array_equal(got["children_ids"], _ids(k["children"])) and np.array_equal(got["loop_edge_ids"], _ids(k["loop_edges"]))
    flat = [c for col in k["grid"] for c in col]
    assert len(got["grid_offsets"]) == len(flat) + 1
    assert np.array_equal(np.diff(got["grid_offsets"]), [len(c) for c in flat])
    assert np.array_equal(got["grid_features"], np.array([f for c in flat for f in c], np.int32))


@pytest.mark.parametrize("seed,kw", [(0, {}), (1, dict(n_kf=2, n_feat=40, n_mp=10, with_origin=False)),
                                     (2, dict(n_kf=5, empty_kf=True)), (3, dict(n_kf=1, n_mp=0)), (4, dict(n_kf=0, n_mp=0))])
def test_load_python_written_archive_and_roundtrip(tmp_path, seed, kw):
    m = make_random_map(seed, **kw)
    blob = serialize_map(m)
    path = tmp_path / "map.bin"
    path.write_bytes(blob)
    ar = orb.MapArchive.load(path)
    info = ar.info()
    assert info["n_mappoints"] == len(m["mappoints"]) and info["n_keyframes"] == len(m["keyframes"])
    assert info["n_origins"] == len(m["origins"]) and info["test_data"] == 0xdeadbeef and info["max_kf_id"] == m["max_kf_id"]
    assert info["total_features"] == sum(k["n"] for k in m["keyframes"])
    assert info["total_observations"] == sum(len(p["obs"]) for p in m["mappoints"])
    # Map::save's second block of map points is what Map::load leaves unread
    assert info["trailing_bytes"] == len(blob) - len(serialize_map(m, second_copy=False))
    for i, k in enumerate(m["keyframes"]):
        _check_keyframe(ar.keyframe(i), k)
    for i, k in enumerate(m["origins"]):
        _check_keyframe(ar.keyframe(i, group=1), k)
    mp = ar.mappoints()
    want = m["mappoints"]
    assert np.array_equal(mp["ids"], np.array([p["id"] for p in want], np.uint64))
    if want:
        assert np.array_equal(mp["world_pos"], np.stack([p["world_pos"][:, 0] for p in want]))
        assert np.array_equal(mp["normal"], np.stack([p["normal"][:, 0] for p in want]))
        assert np.array_equal(mp["desc"], np.concatenate([p["desc"] for p in want]))
    assert np.array_equal(mp["ref_kf"], _ids([p["ref_kf"] for p in want]))
    assert np.array_equal(mp["n_obs"], np.array([p["n_obs"] for p in want], np.int32))
    assert np.array_equal(mp["visible"], np.array([p["visible"] for p in want], np.int32))
    assert np.array_equal(mp["min_dist"], np.array([p["min_dist"] for p in want], np.float32))
    assert np.array_equal(np.diff(mp["obs_offsets"]), [len(p["obs"]) for p in want])
    assert np.array_equal(mp["obs_kf"], _ids([o[0] for p in want for o in p["obs"]]))
    assert np.array_equal(mp["obs_idx"], np.array([-1 if o[0] < 0 else o[1] for p in want for o in p["obs"]], np.int64))
    # the gather of MapPoint::ComputeDistinctiveDescriptors: valid observations of keyframes in the map, in stored order
    desc, off = ar.observed_descriptors()
    by_id = {k["id"]: k for k in m["keyframes"]}
    rows = [[by_id[kf]["desc"][idx] for kf, idx in p["obs"] if kf >= 0] for p in want]
    assert np.array_equal(np.diff(off), [len(r) for r in rows])
    if len(desc):
        assert np.array_equal(desc, np.stack([r for rr in rows for r in rr]))
    # writing it back reproduces the file byte for byte
    out = tmp_path / "again.bin"
    ar.save(out)
    assert out.read_bytes() == blob


def test_builder_writes_the_same_bytes_as_the_python_statement(tmp_path):
    m = make_random_map(7, n_kf=3, n_feat=90, n_mp=40, gba=False)
    ar = orb.MapArchive.create()
    for k in m["keyframes"]:
        info = dict(id=k["id"], frame_id=k["frame_id"], timestamp=k["timestamp"], scale_factor=float(k["scale_factor"]),
                    log_scale_factor=float(k["log_scale_factor"]), fx=k["fx"], fy=k["fy"], cx=k["cx"], cy=k["cy"], bf=k["bf"], b=k["b"],
                    th_depth=k["th_depth"], min_x=k["min_x"], min_y=k["min_y"], max_x=k["max_x"], max_y=k["max_y"],
                    half_baseline=float(k["half_baseline"]), has_parent=int(k["parent"] is not None),
                    parent_id=0 if k["parent"] is None else k["parent"], first_connection=int(k["first_connection"]),
                    not_erase=int(k["not_erase"]))
        kp = np.zeros(k["n"], orb.KP_DTYPE)
        kpu = np.zeros(k["n"], orb.KP_DTYPE)
        for dst, src in ((kp, k["keys"]), (kpu, k["keys_un"])):
            for f in ("x", "y", "angle", "response", "octave", "class_id"):
                dst[f] = [q[f] for q in src]
        ar.add_keyframe(info, kp, kpu, k["desc"], _ids(k["mappoint_ids"]), k["scale_factors"], k["level_sigma2"], k["inv_level_sigma2"],
                        k["Tcw"], k["K"], uright=k["uright"], depth=k["depth"])
    for i, k in enumerate(m["keyframes"]):
        ar.set_keyframe_links(i, [c[0] for c in k["connected"]], [c[1] for c in k["connected"]], k["ordered_ids"], k["ordered_weights"],
                              k["children"], k["loop_edges"])
    ar.add_origin(0)
    for p in m["mappoints"]:
        ar.add_mappoint(p["id"], p["world_pos"], p["normal"], p["desc"], -1 if p["ref_kf"] is None else p["ref_kf"],
                        [o[0] for o in p["obs"]], [o[1] for o in p["obs"]], first_kf=p["first_kf"], visible=p["visible"], found=p["found"],
                        min_dist=float(p["min_dist"]), max_dist=float(p["max_dist"]))
    out = tmp_path / "built.bin"
    ar.save(out)
    assert out.read_bytes() == serialize_map(m)


def test_null_map_pointer_and_errors(tmp_path):
    p = tmp_path / "null.bin"
    p.write_bytes(serialize_map(None))
    assert orb.MapArchive.load(p).info()["n_keyframes"] == 0
    with pytest.raises(OrbError):
        orb.MapArchive.load(tmp_path / "does_not_exist.bin")
    blob = serialize_map(make_random_map(5, n_kf=2, n_feat=30, n_mp=8))
    end_of_load = len(blob) - orb.MapArchive.load(_write(tmp_path, "ok.bin", blob)).info()["trailing_bytes"]
    rng = np.random.default_rng(0)
    for cut in sorted(set([0, 1, 2, 7, 11, 15] + [int(c) for c in rng.integers(0, end_of_load - 1, 60)])):
        with pytest.raises(OrbError):
            orb.MapArchive.load(_write(tmp_path, "cut.bin", blob[:cut]))
    # corrupted bytes must never crash the loader: it either reports an error or returns a map
    for trial in range(200):
        b = bytearray(blob)
        for pos in rng.integers(0, end_of_load, 3):
            b[int(pos)] = int(rng.integers(0, 256))
        try:
            orb.MapArchive.load(_write(tmp_path, "fuzz.bin", bytes(b))).info()
        except OrbError:
            pass
    ar = orb.MapArchive.load(_write(tmp_path, "ok.bin", blob))
    with pytest.raises(OrbError):
        ar.keyframe_info(99)
    with pytest.raises(OrbError):
        ar.keyframe_info(0, group=2)


def _write(tmp_path, name, data):
    p = tmp_path / name
    p.write_bytes(data)
    return p


def test_cpp_map_archive_class(tmp_path):
    """ORB_SLAM2::MapArchiveB200 (orbslam_mapsave_b200/host/MapArchive.h) through the C++ driver: counts, per-keyframe and
    per-map-point contents (as order-sensitive checksums) and the observation gather equal the logical map; Save() reproduces
    the file.  Host-only: runs without a GPU."""
    import subprocess
    drv = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "host_api_driver")
    if not os.path.exists(drv):
        pytest.skip("host_api_driver not built")
    m = make_random_map(21, n_kf=4, n_feat=70, n_mp=35)
    blob = serialize_map(m)
    src, out, dump = tmp_path / "in.bin", tmp_path / "out.bin", tmp_path / "dump.txt"
    src.write_bytes(blob)
    subprocess.check_call([drv, "maparchive", str(src), str(out), str(dump)])
    assert out.read_bytes() == blob
    lines = dump.read_text().split("\n")
    M = (1 << 64) - 1

    def dsum(a):
        s = 0
        for v in np.asarray(a, np.uint8).reshape(-1):
            s = (s * 1000003 + int(v)) & M
        return s

    def lsum(vals):
        s = 0
        for v in vals:
            s = s * 31 + v
            s = ((s + (1 << 63)) & M) - (1 << 63)          # C long wrap-around
        return s

    assert lines[0] == f"map {len(m['keyframes'])} {len(m['mappoints'])} {m['max_kf_id']} 1"
    kfl = [ln.split() for ln in lines if ln.startswith("kf ")]
    assert len(kfl) == len(m["keyframes"])
    for got, k in zip(kfl, m["keyframes"]):
        assert [int(got[1]), int(got[2]), int(got[3]), int(got[4])] == [k["id"], k["frame_id"], k["n"], k["n"]]
        assert int(got[5]) == dsum(k["desc"])
        assert int(got[6]) == lsum([-1 if v < 0 else v for v in k["mappoint_ids"]])
        kx = sum(float(q["x"]) + 2.0 * float(q["y"]) + float(q["angle"]) + q["octave"] for q in k["keys_un"])
        assert abs(float(got[7]) - kx) <= 1e-6 * max(1.0, abs(kx))
        assert np.float32(float(got[8])) == k["Tcw"][1, 3] and np.float32(float(got[9])) == k["K"][0, 2]
        assert int(got[10]) == int(k["parent"] is not None)
        assert [int(v) for v in got[11:16]] == [len(k["connected"]), len(k["ordered_ids"]), len(k["children"]), len(k["loop_edges"]),
                                                sum(len(c) for col in k["grid"] for c in col)]
    mpl = [ln.split() for ln in lines if ln.startswith("mp ")]
    assert len(mpl) == len(m["mappoints"])
    for got, p in zip(mpl, m["mappoints"]):
        assert int(got[1]) == p["id"] and int(got[2]) == dsum(p["desc"])
        assert int(got[3]) == (-1 if p["ref_kf"] is None else p["ref_kf"]) and int(got[4]) == len(p["obs"])
        assert int(got[5]) == lsum([(-1 * 7 + -1) if kf < 0 else kf * 7 + idx for kf, idx in p["obs"]])
        assert np.float32(float(got[6])) == p["world_pos"][2, 0]
    by_id = {k["id"]: k for k in m["keyframes"]}
    rows = [by_id[kf]["desc"][idx] for p in m["mappoints"] for kf, idx in p["obs"] if kf >= 0]
    obs = [ln.split() for ln in lines if ln.startswith("observed ")][0]
    assert int(obs[1]) == len(rows) == int(obs[3]) and int(obs[2]) == dsum(np.stack(rows))
