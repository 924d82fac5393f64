"""ctypes binding of liborb_b200.so (include/orb_b200.h) plus thin numpy/torch conveniences.

This is plumbing only: every computation happens in the CUDA library.  Importing works without a GPU (so the CPU
test-suite can check that the library loads and exports every symbol of the header); any compute call without a CUDA
device raises OrbError — there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORB_B200_LIB") or os.path.join(_HERE, "liborb_b200.so")   # override: A/B experiments with other builds

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
CAND_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("r", "<i4")])

ORB_OK, ORB_ERR_CUDA, ORB_ERR_ARG, ORB_ERR_CAPACITY, ORB_ERR_OVERFLOW, ORB_ERR_GEOMETRY = 0, -1, -2, -3, -4, -5
STAGE_PYRAMID, STAGE_FAST, STAGE_OCTREE, STAGE_BLUR, STAGE_DESCRIBE, STAGE_ALL = 1, 2, 4, 8, 16, 31
TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30

# every symbol include/orb_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "orb_last_error", "orb_device_count",
    "orbx_create", "orbx_destroy", "orbx_tables", "orbx_compute_tables", "orbx_level_size", "orbx_max_keypoints", "orbx_extract",
    "orbx_extract_batch", "orbx_extract_batch_ptrs", "orbx_extract_batch_device", "orbx_host_register", "orbx_host_unregister", "orbx_check_status", "orbx_get_pyramid_level", "orbx_get_pyramid",
    "orbx_get_blurred_level", "orbx_get_candidates", "orbx_launch_count", "orbx_run_stages_device", "orbx_stereo_matches",
    "orbm_descriptor_distance", "orbm_hamming_top2", "orbm_hamming_top2_batch_device", "orbm_allpairs_device", "orbm_allpairs_multi",
    "orbm_search_by_bow_kf_frame", "orbm_search_by_bow_kf_kf", "orbm_search_for_triangulation", "orbm_three_maxima",
    "orbm_search_by_bow_batch", "orbm_search_for_triangulation_batch",
    "orbm_popc_peak", "orbm_distinctive_descriptors", "orb_h2d_probe", "orb_h2d_probe_at",
    "orbm_search_by_projection_map", "orbm_search_by_projection_frame", "orbm_search_by_projection_frame_batch",
    "orbm_search_for_initialization", "orbm_search_windows", "orbm_search_windows_best",
    "orbv_create", "orbv_load_text", "orbv_load_binary", "orbv_save_binary", "orbv_destroy", "orbv_info", "orbv_transform",
    "orbv_transform_device",
    "orbmap_load", "orbmap_save", "orbmap_create", "orbmap_destroy", "orbmap_get_info", "orbmap_keyframe_get_info",
    "orbmap_keyframe_arrays", "orbmap_keyframe_links", "orbmap_keyframe_grid", "orbmap_mappoints", "orbmap_observations", "orbmap_mappoint_record",
    "orbmap_observed_descriptors", "orbmap_add_mappoint", "orbmap_add_keyframe", "orbmap_set_keyframe_links", "orbmap_add_origin",
]


class OrbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"orb_b200 error {code}: {msg}")
        self.code = code


class FeatVecC(C.Structure):
    _fields_ = [("n_nodes", C.c_int), ("node_ids", C.c_void_p), ("offsets", C.c_void_p), ("features", C.c_void_p)]


class ViewC(C.Structure):
    _fields_ = [("n", C.c_int), ("desc", C.c_void_p), ("flag", C.c_void_p), ("angle", C.c_void_p), ("x", C.c_void_p),
                ("y", C.c_void_p), ("octave", C.c_void_p), ("uright", C.c_void_p), ("fv", FeatVecC)]


class GridViewC(C.Structure):
    """orbm_grid_view of include/orb_b200.h"""
    _fields_ = [("n", C.c_int), ("desc", C.c_void_p), ("x", C.c_void_p), ("y", C.c_void_p), ("octave", C.c_void_p),
                ("angle", C.c_void_p), ("uright", C.c_void_p), ("blocked", C.c_void_p), ("grid_cols", C.c_int), ("grid_rows", C.c_int),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float), ("inv_w", C.c_float),
                ("inv_h", C.c_float), ("cell_offsets", C.c_void_p), ("cell_features", C.c_void_p), ("scale_factors", C.c_void_p),
                ("n_levels", C.c_int)]


class FrameSearchJobC(C.Structure):
    """orbm_frame_search_job of include/orb_b200.h"""
    _fields_ = [("cur", C.POINTER(GridViewC)), ("Tcw_cur", C.c_void_p), ("Tcw_last", C.c_void_p), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("mbf", C.c_float), ("mb", C.c_float), ("n_last", C.c_int), ("has_point", C.c_void_p),
                ("world", C.c_void_p), ("octave", C.c_void_p), ("angle", C.c_void_p), ("desc", C.c_void_p), ("claims", C.c_void_p),
                ("th", C.c_float), ("mono", C.c_int), ("owner", C.c_void_p), ("n_matches", C.c_int)]


class MapInfoC(C.Structure):
    """orbmap_info of include/orb_b200.h"""
    _fields_ = [("n_mappoints", C.c_int32), ("n_keyframes", C.c_int32), ("n_origins", C.c_int32), ("test_data", C.c_uint32),
                ("max_kf_id", C.c_uint64), ("total_features", C.c_int64), ("total_observations", C.c_int64),
                ("trailing_bytes", C.c_int64)]


class MapKeyFrameInfoC(C.Structure):
    """orbmap_keyframe_info of include/orb_b200.h"""
    _fields_ = ([(n, C.c_uint64) for n in ("id", "frame_id", "next_id", "parent_id")] + [("timestamp", C.c_double)] +
                [(n, C.c_int32) for n in ("n", "n_keys", "n_keys_un", "n_uright", "n_depth", "desc_rows", "desc_cols",
                                          "n_mappoint_slots", "n_levels", "n_scale_factors", "grid_cols", "grid_rows", "min_x", "min_y",
                                          "max_x", "max_y", "n_connected", "n_ordered", "n_children", "n_loop_edges", "has_parent",
                                          "is_bad", "not_erase", "to_be_erased", "first_connection")] +
                [(n, C.c_float) for n in ("scale_factor", "log_scale_factor", "fx", "fy", "cx", "cy", "invfx", "invfy", "bf", "b",
                                          "th_depth", "grid_inv_w", "grid_inv_h", "half_baseline")])


_lib = None


def lib():
    """Load the CUDA library; raise loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OrbError(ORB_ERR_CUDA, f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                     "(make -C orbslam_mapsave_b200/csrc); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    L.orb_last_error.restype = C.c_char_p
    L.orb_device_count.restype = i32
    L.orbx_create.restype = i32
    L.orbx_create.argtypes = [C.POINTER(vp), i32, f32, i32, i32, i32, i32, i32, i32, i32]
    L.orbx_destroy.restype = None
    L.orbx_destroy.argtypes = [vp]
    L.orbx_tables.restype = i32
    L.orbx_tables.argtypes = [vp] * 6
    L.orbx_compute_tables.restype = i32
    L.orbx_compute_tables.argtypes = [i32, f32, i32, vp, vp, vp, vp, vp]
    L.orbx_level_size.restype = i32
    L.orbx_level_size.argtypes = [vp, i32, vp, vp]
    L.orbx_max_keypoints.restype = i32
    L.orbx_max_keypoints.argtypes = [vp]
    L.orbx_extract.restype = i32
    L.orbx_extract.argtypes = [vp, vp, i32, i32, i32, vp, i32, vp, vp, i32, vp]
    L.orbx_extract_batch.restype = i32
    L.orbx_extract_batch.argtypes = [vp, vp, i32, i32, i32, i32, sz, vp, i32, sz, vp, vp, i32, vp]
    L.orbx_extract_batch_ptrs.restype = i32
    L.orbx_extract_batch_ptrs.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, i32, vp]
    L.orbx_host_register.restype = i32
    L.orbx_host_register.argtypes = [vp, sz]
    L.orbx_host_unregister.restype = i32
    L.orbx_host_unregister.argtypes = [vp]
    L.orbx_extract_batch_device.restype = i32
    L.orbx_extract_batch_device.argtypes = [vp, vp, i32, vp, vp, vp, i32, vp, vp]
    L.orbx_check_status.restype = i32
    L.orbx_check_status.argtypes = [vp]
    L.orbx_get_pyramid_level.restype = i32
    L.orbx_get_pyramid_level.argtypes = [vp, i32, i32, i32, vp, i32]
    L.orbx_get_pyramid.restype = i32
    L.orbx_get_pyramid.argtypes = [vp, i32, i32, vp, vp]
    L.orbx_get_blurred_level.restype = i32
    L.orbx_get_blurred_level.argtypes = [vp, i32, i32, vp, i32]
    L.orbx_get_candidates.restype = i32
    L.orbx_get_candidates.argtypes = [vp, i32, i32, vp, i32, vp]
    L.orbx_launch_count.restype = C.c_longlong
    L.orbx_launch_count.argtypes = [vp]
    L.orbx_run_stages_device.restype = i32
    L.orbx_run_stages_device.argtypes = [vp, vp, i32, vp, vp, vp, i32, vp, i32, vp]
    L.orbx_stereo_matches.restype = i32
    L.orbx_stereo_matches.argtypes = [vp, vp, i32, i32, vp, vp, i32, vp, vp, i32, f32, f32, vp, vp]
    L.orbm_descriptor_distance.restype = i32
    L.orbm_descriptor_distance.argtypes = [vp, vp, i32, vp, i32]
    L.orbm_hamming_top2.restype = i32
    L.orbm_hamming_top2.argtypes = [vp, i32, vp, i32, vp, vp, vp, i32]
    L.orbm_hamming_top2_batch_device.restype = i32
    L.orbm_hamming_top2_batch_device.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp]
    L.orbm_allpairs_device.restype = i32
    L.orbm_allpairs_device.argtypes = [vp, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp]
    L.orbm_allpairs_multi.restype = i32
    L.orbm_allpairs_multi.argtypes = [vp, i32, i32, i32, f32, vp, i32, vp, vp, vp]
    L.orbm_search_by_bow_kf_frame.restype = i32
    L.orbm_search_by_bow_kf_frame.argtypes = [C.POINTER(ViewC), C.POINTER(ViewC), f32, i32, vp, vp, i32]
    L.orbm_search_by_bow_kf_kf.restype = i32
    L.orbm_search_by_bow_kf_kf.argtypes = [C.POINTER(ViewC), C.POINTER(ViewC), f32, i32, vp, vp, i32]
    L.orbm_search_for_triangulation.restype = i32
    L.orbm_search_for_triangulation.argtypes = [C.POINTER(ViewC), C.POINTER(ViewC), vp, f32, f32, vp, vp, i32, i32, i32,
                                                vp, vp, vp, i32]
    L.orbm_search_by_bow_batch.restype = i32
    L.orbm_search_by_bow_batch.argtypes = [C.POINTER(ViewC), C.POINTER(ViewC), i32, i32, f32, i32, vp, vp, i32]
    L.orbm_search_for_triangulation_batch.restype = i32
    L.orbm_search_for_triangulation_batch.argtypes = [C.POINTER(ViewC), C.POINTER(ViewC), i32, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, i32]
    L.orbm_three_maxima.restype = i32
    L.orbm_three_maxima.argtypes = [vp, i32, vp, i32]
    L.orbm_popc_peak.restype = i32
    L.orbm_popc_peak.argtypes = [i32, vp, vp]
    L.orb_h2d_probe.restype = i32
    L.orb_h2d_probe.argtypes = [i32, vp, sz, sz, i32, i32, vp, vp, vp]
    L.orb_h2d_probe_at.restype = i32
    L.orb_h2d_probe_at.argtypes = [i32, vp, sz, sz, i32, i32, C.c_longlong, vp, vp, vp, vp]
    L.orbm_distinctive_descriptors.restype = i32
    L.orbm_distinctive_descriptors.argtypes = [vp, vp, i32, vp, i32]
    L.orbm_search_by_projection_map.restype = i32
    L.orbm_search_by_projection_map.argtypes = [C.POINTER(GridViewC), i32, vp, vp, vp, vp, vp, vp, vp, vp, f32, f32, vp, vp, i32]
    L.orbm_search_by_projection_frame.restype = i32
    L.orbm_search_by_projection_frame.argtypes = [C.POINTER(GridViewC), vp, vp, f32, f32, f32, f32, f32, f32, i32, vp, vp, vp, vp, vp,
                                                  vp, f32, i32, i32, vp, vp, i32]
    L.orbm_search_by_projection_frame_batch.restype = i32
    L.orbm_search_by_projection_frame_batch.argtypes = [vp, i32, i32, i32]
    L.orbm_search_for_initialization.restype = i32
    L.orbm_search_for_initialization.argtypes = [C.POINTER(GridViewC), i32, vp, vp, vp, vp, i32, f32, i32, vp, vp, i32]
    L.orbm_search_windows.restype = i32
    L.orbm_search_windows.argtypes = [C.POINTER(GridViewC), i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, i32]
    L.orbm_search_windows_best.restype = i32
    L.orbm_search_windows_best.argtypes = [C.POINTER(GridViewC), i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, i32]
    L.orbv_create.restype = i32
    L.orbv_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i32, vp, vp, vp, vp, i32]
    L.orbv_load_text.restype = i32
    L.orbv_load_text.argtypes = [C.POINTER(vp), C.c_char_p, i32]
    L.orbv_load_binary.restype = i32
    L.orbv_load_binary.argtypes = [C.POINTER(vp), C.c_char_p, i32]
    L.orbv_save_binary.restype = i32
    L.orbv_save_binary.argtypes = [vp, C.c_char_p]
    L.orbv_destroy.restype = None
    L.orbv_destroy.argtypes = [vp]
    L.orbv_info.restype = i32
    L.orbv_info.argtypes = [vp] * 7
    L.orbv_transform.restype = i32
    L.orbv_transform.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    L.orbv_transform_device.restype = i32
    L.orbv_transform_device.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp]
    i64, u64 = C.c_int64, C.c_uint64
    L.orbmap_load.restype = i32
    L.orbmap_load.argtypes = [C.POINTER(vp), C.c_char_p]
    L.orbmap_save.restype = i32
    L.orbmap_save.argtypes = [vp, C.c_char_p]
    L.orbmap_create.restype = i32
    L.orbmap_create.argtypes = [C.POINTER(vp)]
    L.orbmap_destroy.restype = None
    L.orbmap_destroy.argtypes = [vp]
    L.orbmap_get_info.restype = i32
    L.orbmap_get_info.argtypes = [vp, C.POINTER(MapInfoC)]
    L.orbmap_keyframe_get_info.restype = i32
    L.orbmap_keyframe_get_info.argtypes = [vp, i32, i32, C.POINTER(MapKeyFrameInfoC)]
    L.orbmap_keyframe_arrays.restype = i32
    L.orbmap_keyframe_arrays.argtypes = [vp, i32, i32] + [vp] * 11
    L.orbmap_keyframe_links.restype = i32
    L.orbmap_keyframe_links.argtypes = [vp, i32, i32] + [vp] * 6
    L.orbmap_keyframe_grid.restype = i32
    L.orbmap_keyframe_grid.argtypes = [vp, i32, i32, vp, vp, i32, vp, vp]
    L.orbmap_mappoints.restype = i32
    L.orbmap_mappoints.argtypes = [vp] * 13
    L.orbmap_observations.restype = i32
    L.orbmap_observations.argtypes = [vp, vp, vp]
    L.orbmap_mappoint_record.restype = i32
    L.orbmap_mappoint_record.argtypes = [vp, i32, vp, i64, vp]
    L.orbmap_observed_descriptors.restype = i32
    L.orbmap_observed_descriptors.argtypes = [vp, vp, vp, i64, vp]
    L.orbmap_add_mappoint.restype = i32
    L.orbmap_add_mappoint.argtypes = [vp, u64, i64, vp, vp, vp, i64, i32, vp, vp, i32, i32, f32, f32]
    L.orbmap_add_keyframe.restype = i32
    L.orbmap_add_keyframe.argtypes = [vp, C.POINTER(MapKeyFrameInfoC)] + [vp] * 11
    L.orbmap_set_keyframe_links.restype = i32
    L.orbmap_set_keyframe_links.argtypes = [vp, i32, i32, vp, vp, i32, vp, vp, i32, vp, i32, vp]
    L.orbmap_add_origin.restype = i32
    L.orbmap_add_origin.argtypes = [vp, i32]
    _lib = L
    return L


def check(rc):
    if rc != ORB_OK:
        raise OrbError(rc, lib().orb_last_error().decode("utf-8", "replace"))


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a.data_ptr()))          # torch tensor


def device_count():
    return lib().orb_device_count()
