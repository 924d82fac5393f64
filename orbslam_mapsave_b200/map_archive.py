"""The fork's saved map (System::SaveMap / LoadMap, src/System.cc:552-574) as a descriptor source: ctypes binding of the
orbmap_* entry points of include/orb_b200.h.  Parsing happens in the C library (csrc/orb_map.cpp); this class only hands out
numpy views of the flat tables and feeds them to the GPU entry points."""
import ctypes as C

import numpy as np

from . import capi
from .capi import KP_DTYPE, MapInfoC, MapKeyFrameInfoC

_INFO_FIELDS = [f for f, _ in MapKeyFrameInfoC._fields_]


class MapArchive:
    def __init__(self, handle):
        self._h = handle

    @classmethod
    def load(cls, path):
        h = C.c_void_p()
        capi.check(capi.lib().orbmap_load(C.byref(h), str(path).encode()))
        return cls(h)

    @classmethod
    def create(cls):
        h = C.c_void_p()
        capi.check(capi.lib().orbmap_create(C.byref(h)))
        return cls(h)

    def save(self, path):
        capi.check(capi.lib().orbmap_save(self._h, str(path).encode()))

    def close(self):
        if self._h:
            capi.lib().orbmap_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        o = MapInfoC()
        capi.check(capi.lib().orbmap_get_info(self._h, C.byref(o)))
        return {f: getattr(o, f) for f, _ in MapInfoC._fields_}

    def keyframe_info(self, i, group=0):
        o = MapKeyFrameInfoC()
        capi.check(capi.lib().orbmap_keyframe_get_info(self._h, group, i, C.byref(o)))
        return {f: getattr(o, f) for f in _INFO_FIELDS}

    def keyframe(self, i, group=0):
        """Everything stored for one keyframe as numpy arrays (mappoint ids: -1 = no map point)."""
        inf = self.keyframe_info(i, group)
        out = dict(info=inf,
                   keys=np.zeros(inf["n_keys"], KP_DTYPE), keys_un=np.zeros(inf["n_keys_un"], KP_DTYPE),
                   uright=np.zeros(inf["n_uright"], np.float32), depth=np.zeros(inf["n_depth"], np.float32),
                   desc=np.zeros((inf["desc_rows"], max(inf["desc_cols"], 0)), np.uint8),
                   mappoint_ids=np.zeros(inf["n_mappoint_slots"], np.int64),
                   scale_factors=np.zeros(inf["n_scale_factors"], np.float32), level_sigma2=np.zeros(inf["n_scale_factors"], np.float32),
                   inv_level_sigma2=np.zeros(inf["n_scale_factors"], np.float32), Tcw=np.zeros((4, 4), np.float32),
                   K=np.zeros((3, 3), np.float32))
        capi.check(capi.lib().orbmap_keyframe_arrays(self._h, group, i, *[capi._p(out[k]) for k in (
            "keys", "keys_un", "uright", "depth", "desc", "mappoint_ids", "scale_factors", "level_sigma2", "inv_level_sigma2", "Tcw", "K")]))
        links = dict(connected_ids=np.zeros(inf["n_connected"], np.int64), connected_weights=np.zeros(inf["n_connected"], np.int32),
                     ordered_ids=np.zeros(inf["n_ordered"], np.int64), ordered_weights=np.zeros(inf["n_ordered"], np.int32),
                     children_ids=np.zeros(inf["n_children"], np.int64), loop_edge_ids=np.zeros(inf["n_loop_edges"], np.int64))
        capi.check(capi.lib().orbmap_keyframe_links(self._h, group, i, *[capi._p(v) for v in links.values()]))
        out.update(links)
        nc, ne = C.c_int32(), C.c_int32()
        capi.check(capi.lib().orbmap_keyframe_grid(self._h, group, i, None, None, 0, C.byref(nc), C.byref(ne)))
        off, feat = np.zeros(nc.value + 1, np.int32), np.zeros(max(ne.value, 1), np.int32)
        capi.check(capi.lib().orbmap_keyframe_grid(self._h, group, i, capi._p(off), capi._p(feat), ne.value, None, None))
        out["grid_offsets"], out["grid_features"] = off, feat[:ne.value]
        return out

    def mappoints(self):
        n = self.info()["n_mappoints"]
        out = dict(ids=np.zeros(n, np.uint64), world_pos=np.zeros((n, 3), np.float32), normal=np.zeros((n, 3), np.float32),
                   desc=np.zeros((n, 32), np.uint8), ref_kf=np.zeros(n, np.int64), bad=np.zeros(n, np.uint8), n_obs=np.zeros(n, np.int32),
                   visible=np.zeros(n, np.int32), found=np.zeros(n, np.int32), min_dist=np.zeros(n, np.float32),
                   max_dist=np.zeros(n, np.float32), obs_offsets=np.zeros(n + 1, np.int32))
        capi.check(capi.lib().orbmap_mappoints(self._h, *[capi._p(v) for v in out.values()]))
        tot = int(out["obs_offsets"][-1])
        out["obs_kf"], out["obs_idx"] = np.zeros(tot, np.int64), np.zeros(tot, np.int64)
        capi.check(capi.lib().orbmap_observations(self._h, capi._p(out["obs_kf"]), capi._p(out["obs_idx"])))
        return out

    def observed_descriptors(self):
        """(desc[total, 32], offsets[n_mappoints + 1]): the ragged batch orbm_distinctive_descriptors consumes."""
        n = self.info()["n_mappoints"]
        off = np.zeros(n + 1, np.int32)
        tot = C.c_int64()
        capi.check(capi.lib().orbmap_observed_descriptors(self._h, None, capi._p(off), 0, C.byref(tot)))
        desc = np.zeros((max(tot.value, 1), 32), np.uint8)
        capi.check(capi.lib().orbmap_observed_descriptors(self._h, capi._p(desc), capi._p(off), tot.value, None))
        return desc[:tot.value], off

    def distinctive_descriptors(self, device=0):
        """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:483-548) for every map point of the archive, on the GPU:
        returns (descriptor[n, 32], chosen index inside the observation set or -1)."""
        from .matcher import distinctive_descriptors
        desc, off = self.observed_descriptors()
        best = distinctive_descriptors(desc, off, device=device)
        out = np.zeros((len(best), 32), np.uint8)
        ok = best >= 0
        out[ok] = desc[off[:-1][ok] + best[ok]]
        return out, best

    # ---- building ----------------------------------------------------------------------------------------------------
    def add_mappoint(self, id, world_pos, normal, desc, ref_kf=-1, obs_kf=(), obs_idx=(), first_kf=0, visible=1, found=1,
                     min_dist=0.0, max_dist=0.0):
        wp, nv = np.ascontiguousarray(world_pos, np.float32), np.ascontiguousarray(normal, np.float32)
        d = np.ascontiguousarray(desc, np.uint8)
        ok, oi = np.ascontiguousarray(obs_kf, np.int64), np.ascontiguousarray(obs_idx, np.int64)
        capi.check(capi.lib().orbmap_add_mappoint(self._h, int(id), int(first_kf), capi._p(wp), capi._p(nv), capi._p(d), int(ref_kf),
                                                  len(ok), capi._p(ok), capi._p(oi), visible, found, min_dist, max_dist))

    def add_keyframe(self, info, keys, keys_un, desc, mappoint_ids, scale_factors, level_sigma2, inv_level_sigma2, Tcw, K,
                     uright=None, depth=None):
        o = MapKeyFrameInfoC()
        for k, v in info.items():
            setattr(o, k, v)
        o.n = len(keys)
        o.n_levels = len(scale_factors)
        arrs = [np.ascontiguousarray(keys, KP_DTYPE), np.ascontiguousarray(keys_un, KP_DTYPE),
                None if uright is None else np.ascontiguousarray(uright, np.float32),
                None if depth is None else np.ascontiguousarray(depth, np.float32), np.ascontiguousarray(desc, np.uint8),
                np.ascontiguousarray(mappoint_ids, np.int64), np.ascontiguousarray(scale_factors, np.float32),
                np.ascontiguousarray(level_sigma2, np.float32), np.ascontiguousarray(inv_level_sigma2, np.float32),
                np.ascontiguousarray(Tcw, np.float32), np.ascontiguousarray(K, np.float32)]
        capi.check(capi.lib().orbmap_add_keyframe(self._h, C.byref(o), *[capi._p(a) for a in arrs]))

    def set_keyframe_links(self, i, connected_ids=(), connected_weights=(), ordered_ids=(), ordered_weights=(), children_ids=(),
                           loop_edge_ids=()):
        a = [np.ascontiguousarray(connected_ids, np.int64), np.ascontiguousarray(connected_weights, np.int32),
             np.ascontiguousarray(ordered_ids, np.int64), np.ascontiguousarray(ordered_weights, np.int32),
             np.ascontiguousarray(children_ids, np.int64), np.ascontiguousarray(loop_edge_ids, np.int64)]
        capi.check(capi.lib().orbmap_set_keyframe_links(self._h, i, len(a[0]), capi._p(a[0]), capi._p(a[1]), len(a[2]), capi._p(a[2]),
                                                        capi._p(a[3]), len(a[4]), capi._p(a[4]), len(a[5]), capi._p(a[5])))

    def add_origin(self, keyframe_index):
        capi.check(capi.lib().orbmap_add_origin(self._h, keyframe_index))
