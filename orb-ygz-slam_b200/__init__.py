"""orb_ygz_slam_b200 — Python binding (ctypes) of libygz_b200.so, the B200-native ORB-YGZ-SLAM image
front end (ORBextractor / ORBmatcher Hamming search / SparseImgAlign).

The compute path is hand-written CUDA for sm_100a behind the C ABI of include/ygz_b200.h; this module
only marshals numpy arrays into it.  There is NO CPU fallback: importing works without a GPU (so the
symbol table can be checked), but creating any handle without the extension or without a CUDA device
raises.  The directory name has a hyphen (task layout), so load it through `__graft_entry__.load_package()`
or `importlib` with the module name `orb_ygz_slam_b200`.
"""
import ctypes as C
import os
import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libygz_b200.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])

u8p = C.POINTER(C.c_uint8)
f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)
u32p = C.POINTER(C.c_uint32)


class YgzError(RuntimeError):
    pass


class OrbConfig(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int),
                ("ini_th_fast", C.c_int), ("min_th_fast", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("max_batch", C.c_int), ("device", C.c_int)]


class FrameView(C.Structure):
    _fields_ = [("n", C.c_int), ("keys", C.c_void_p), ("desc", u8p), ("u_right", f32p), ("outlier", u8p),
                ("map_point", i32p),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("mb", C.c_float), ("mbf", C.c_float),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float),
                ("n_levels", C.c_int), ("scale_factors", f32p), ("inv_scale_factors", f32p),
                ("log_scale_factor", C.c_float), ("Rcw", C.c_float * 9), ("tcw", C.c_float * 3),
                ("fv_nodes", C.c_int), ("fv_node_id", u32p), ("fv_offset", i32p), ("fv_index", u32p)]


class MapPoints(C.Structure):
    _fields_ = [("m", C.c_int), ("world_pos", f32p), ("desc", u8p), ("bad", u8p), ("observations", i32p),
                ("min_dist", f32p), ("max_dist", f32p), ("track_in_view", u8p), ("track_proj_x", f32p),
                ("track_proj_y", f32p), ("track_proj_xr", f32p), ("track_scale_level", i32p),
                ("track_view_cos", f32p), ("normal", f32p)]


class SiaProblem(C.Structure):
    _fields_ = [("n_levels", C.c_int), ("ref_levels", C.POINTER(u8p)), ("cur_levels", C.POINTER(u8p)),
                ("level_w", i32p), ("level_h", i32p), ("inv_scale_factors", f32p),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("n", C.c_int), ("ref_px", f32p), ("world_xyz", f32p), ("valid", u8p),
                ("ref_q", C.c_float * 4), ("ref_t", C.c_float * 3),
                ("cur_q", C.c_float * 4), ("cur_t", C.c_float * 3)]


class DirectFrame(C.Structure):
    """== ygz_direct_frame (include/ygz_b200.h)."""
    _fields_ = [("level", u8p * 16), ("w", C.c_int * 16), ("h", C.c_int * 16), ("pitch", C.c_int * 16),
                ("q", C.c_float * 4), ("t", C.c_float * 3)]


class DirectItem(C.Structure):
    """== ygz_direct_item."""
    _fields_ = [("ref", C.c_int), ("px_ref", C.c_float * 2), ("octave", C.c_int), ("pw", C.c_float * 3),
                ("px_cur", C.c_float * 2)]


DIRECT_ITEM_DTYPE = np.dtype([("ref", np.int32), ("px_ref", np.float32, 2), ("octave", np.int32), ("pw", np.float32, 3),
                              ("px_cur", np.float32, 2)])


class Counters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("frames", "keypoints", "matches", "sia_iters", "bytes_in", "bytes_out",
                                          "kernel_launches", "elapsed_ns")]


class ProjectionCamera(C.Structure):      # ygz_projection_camera
    _fields_ = [(k, C.c_float) for k in ("fx", "fy", "cx", "cy", "mb", "mbf", "min_x", "min_y", "max_x", "max_y")]


# every symbol include/ygz_b200.h declares (tests/test_abi.py checks the .so exports all of them)
ABI_SYMBOLS = [
    "ygz_last_error", "ygz_device_count", "ygz_device_info",
    "ygz_orb_create", "ygz_orb_destroy", "ygz_orb_tables", "ygz_orb_level_size", "ygz_orb_max_keypoints", "ygz_orb_max_batch",
    "ygz_orb_extract", "ygz_orb_extract_async", "ygz_orb_wait", "ygz_orb_extract_frame", "ygz_orb_extract_frame_dso", "ygz_orb_extract_frame_dso_batch",
    "ygz_orb_compute_pyramid", "ygz_orb_get_level", "ygz_orb_get_blurred_level",
    "ygz_orb_upload", "ygz_orb_run", "ygz_orb_download", "ygz_orb_sync", "ygz_orb_get_candidates",
    "ygz_orb_time", "ygz_debug_cos_sin", "ygz_orb_device_outputs", "ygz_orb_device_level", "ygz_orb_stream", "ygz_orb_counters",
    "ygz_descriptor_distance", "ygz_matcher_create", "ygz_matcher_destroy", "ygz_hamming_bf_top2",
    "ygz_hamming_pairs", "ygz_search_by_bow", "ygz_search_by_bow_kf", "ygz_search_for_triangulation", "ygz_search_by_projection_points",
    "ygz_search_by_projection_last", "ygz_search_by_projection_kf", "ygz_search_for_initialization",
    "ygz_matcher_set_vocabulary", "ygz_search_by_bow_batch", "ygz_search_by_bow_batch_async", "ygz_search_by_bow_batch_time", "ygz_frontend_time", "ygz_frontend_time2", "ygz_frontend_time_n",
    "ygz_sia_create", "ygz_sia_destroy", "ygz_sia_run", "ygz_sia_run_orb", "ygz_sia_time", "ygz_sia_last_trace",
    "ygz_direct_create", "ygz_direct_destroy", "ygz_direct_upload_frame", "ygz_direct_set_pose",
    "ygz_find_direct_projection", "ygz_direct_time", "ygz_compute_stereo_matches",
    "ygz_stereo_batch", "ygz_stereo_batch_async", "ygz_stereo_frontend_time",
    "ygz_search_by_projection_sim3", "ygz_fuse_sim3", "ygz_search_by_sim3", "ygz_fuse_candidates",
    "ygz_search_by_projection_last_batch", "ygz_search_by_projection_points_batch", "ygz_search_by_projection_batch_time",
    "ygz_vocab_load", "ygz_vocab_from_memory", "ygz_vocab_destroy", "ygz_vocab_info", "ygz_vocab_transform", "ygz_vocab_transform_orb",
]

_lib = None


def _ptr(a, t):
    if a is None:
        return C.cast(None, t)
    return a.ctypes.data_as(t)


def lib():
    """Load libygz_b200.so.  Raises (loudly) when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise YgzError("libygz_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a).  There is no CPU fallback for the product path.")
    L = C.CDLL(LIB_PATH)
    sigs = {}
    sigs.setdefault('ygz_last_error', {})['restype'] = C.c_char_p
    sigs.setdefault('ygz_orb_create', {})['argtypes'] = [C.POINTER(OrbConfig), C.POINTER(C.c_void_p)]
    sigs.setdefault('ygz_orb_destroy', {})['argtypes'] = [C.c_void_p]
    sigs.setdefault('ygz_orb_tables', {})['argtypes'] = [C.c_void_p, f32p, f32p, f32p, f32p, i32p]
    sigs.setdefault('ygz_orb_level_size', {})['argtypes'] = [C.c_void_p, C.c_int, i32p, i32p]
    sigs.setdefault('ygz_orb_max_keypoints', {})['argtypes'] = [C.c_void_p]
    sigs.setdefault('ygz_orb_extract', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    sigs.setdefault('ygz_orb_extract_frame', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, i32p, C.c_int]
    sigs.setdefault('ygz_orb_extract_frame_dso_batch', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int, i32p, C.c_void_p, C.c_void_p, i32p, C.c_int, i32p]
    sigs.setdefault('ygz_orb_extract_frame_dso', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, i32p, C.c_int, i32p]
    sigs.setdefault('ygz_orb_compute_pyramid', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int]
    sigs.setdefault('ygz_orb_get_level', {})['argtypes'] = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    sigs.setdefault('ygz_orb_get_blurred_level', {})['argtypes'] = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    sigs.setdefault('ygz_orb_upload', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int]
    sigs.setdefault('ygz_orb_run', {})['argtypes'] = [C.c_void_p, C.c_int]
    sigs.setdefault('ygz_orb_download', {})['argtypes'] = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    sigs.setdefault('ygz_orb_sync', {})['argtypes'] = [C.c_void_p]
    sigs.setdefault('ygz_orb_get_candidates', {})['argtypes'] = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    sigs.setdefault('ygz_orb_time', {})['argtypes'] = [C.c_void_p, C.c_int, C.c_int, f32p, i32p]
    sigs.setdefault('ygz_orb_device_outputs', {})['argtypes'] = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), i32p]
    sigs.setdefault('ygz_orb_device_level', {})['argtypes'] = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), i32p, C.POINTER(C.c_size_t)]
    sigs.setdefault('ygz_orb_stream', {})['restype'] = C.c_void_p
    sigs.setdefault('ygz_orb_stream', {})['argtypes'] = [C.c_void_p]
    sigs.setdefault('ygz_orb_counters', {})['argtypes'] = [C.c_void_p, C.POINTER(Counters), C.c_int]
    sigs.setdefault('ygz_device_info', {})['argtypes'] = [C.c_int, C.c_char_p, i32p, C.POINTER(C.c_size_t)]
    sigs.setdefault('ygz_descriptor_distance', {})['argtypes'] = [C.c_void_p, C.c_void_p]
    sigs.setdefault('ygz_matcher_create', {})['argtypes'] = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    sigs.setdefault('ygz_matcher_destroy', {})['argtypes'] = [C.c_void_p]
    sigs.setdefault('ygz_hamming_bf_top2', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, i32p, i32p, i32p]
    sigs.setdefault('ygz_hamming_pairs', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, i32p, i32p, C.c_int, i32p]
    sigs.setdefault('ygz_search_by_bow', {})['argtypes'] = [C.c_void_p, C.c_float, C.c_int, C.POINTER(FrameView), C.POINTER(MapPoints), C.POINTER(FrameView), i32p]
    sigs.setdefault('ygz_search_by_bow_kf', {})['argtypes'] = [C.c_void_p, C.c_float, C.c_int, C.POINTER(FrameView), C.POINTER(FrameView), C.POINTER(MapPoints), i32p]
    sigs.setdefault('ygz_search_for_triangulation', {})['argtypes'] = [C.c_void_p, C.c_int, C.POINTER(FrameView), C.POINTER(FrameView), f32p, C.c_float, C.c_float, f32p, C.c_int, i32p]
    sigs.setdefault('ygz_search_by_projection_points', {})['argtypes'] = [C.c_void_p, C.c_float, C.POINTER(FrameView), C.POINTER(MapPoints), i32p, C.c_int, C.c_float, C.c_int]
    sigs.setdefault('ygz_search_by_projection_last', {})['argtypes'] = [C.c_void_p, C.c_int, C.POINTER(FrameView), C.POINTER(FrameView), C.POINTER(MapPoints), C.c_float, C.c_int, C.c_int]
    sigs.setdefault('ygz_search_by_projection_kf', {})['argtypes'] = [C.c_void_p, C.c_int, C.POINTER(FrameView), C.POINTER(FrameView), C.POINTER(MapPoints), u8p, C.c_float, C.c_int]
    sigs.setdefault('ygz_search_for_initialization', {})['argtypes'] = [C.c_void_p, C.c_float, C.c_int, C.POINTER(FrameView), C.POINTER(FrameView), f32p, i32p, C.c_int]
    sigs.setdefault('ygz_search_by_bow_batch', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_float, C.c_int, i32p, i32p, C.c_int, u32p, i32p, i32p]
    sigs.setdefault('ygz_compute_stereo_matches', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, f32p, f32p]
    sigs.setdefault('ygz_search_by_projection_sim3', {})['argtypes'] = [C.c_void_p, C.POINTER(FrameView), f32p, C.POINTER(MapPoints), i32p, C.c_int, i32p, C.c_int]
    sigs.setdefault('ygz_fuse_sim3', {})['argtypes'] = [C.c_void_p, C.POINTER(FrameView), f32p, C.POINTER(MapPoints), i32p, C.c_int, C.c_float, i32p]
    sigs.setdefault('ygz_fuse_candidates', {})['argtypes'] = [C.c_void_p, C.POINTER(FrameView), f32p, f32p, C.POINTER(MapPoints), i32p, C.c_int, C.c_float, i32p, i32p]
    sigs.setdefault('ygz_search_by_sim3', {})['argtypes'] = [C.c_void_p, C.POINTER(FrameView), C.POINTER(FrameView), C.POINTER(MapPoints), i32p, i32p, C.c_float, f32p, f32p, C.c_float]
    sigs.setdefault('ygz_search_by_projection_last_batch', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.POINTER(ProjectionCamera), C.c_int, i32p, i32p, f32p, u8p, f32p, u8p, f32p,
                                                                            C.c_float, C.c_int, C.c_int, C.c_int, i32p, i32p, i32p]
    sigs.setdefault('ygz_search_by_projection_points_batch', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.POINTER(ProjectionCamera), C.c_int, i32p, C.c_int, f32p, f32p, f32p, f32p,
                                                                              i32p, u8p, u8p, u8p, f32p, C.c_float, C.c_float, C.c_int, i32p, i32p, i32p]
    sigs.setdefault('ygz_search_by_projection_batch_time', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, f32p]
    sigs.setdefault('ygz_stereo_batch', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, f32p, f32p, i32p, i32p, i32p, i32p]
    sigs.setdefault('ygz_stereo_batch_async', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    sigs.setdefault('ygz_stereo_frontend_time', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, f32p, f32p]
    sigs.setdefault('ygz_vocab_load', {})['argtypes'] = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    sigs.setdefault('ygz_vocab_from_memory', {})['argtypes'] = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
    sigs.setdefault('ygz_vocab_destroy', {})['argtypes'] = [C.c_void_p]
    sigs.setdefault('ygz_vocab_info', {})['argtypes'] = [C.c_void_p, i32p, i32p, i32p, i32p, i32p, i32p]
    sigs.setdefault('ygz_vocab_transform', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, i32p, f32p, i32p]
    sigs.setdefault('ygz_vocab_transform_orb', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, i32p, f32p, i32p, f32p]
    sigs.setdefault('ygz_direct_create', {})['argtypes'] = [C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]
    sigs.setdefault('ygz_direct_destroy', {})['argtypes'] = [C.c_void_p]
    sigs.setdefault('ygz_direct_upload_frame', {})['argtypes'] = [C.c_void_p, C.c_int, C.POINTER(DirectFrame), C.c_int]
    sigs.setdefault('ygz_direct_set_pose', {})['argtypes'] = [C.c_void_p, C.c_int, f32p, f32p]
    sigs.setdefault('ygz_find_direct_projection', {})['argtypes'] = [C.c_void_p, C.c_int, f32p, f32p, f32p, C.c_float, C.c_int, C.c_void_p, C.c_int, f32p, i32p, u8p]
    sigs.setdefault('ygz_direct_time', {})['argtypes'] = [C.c_void_p, C.c_int, f32p]
    sigs.setdefault('ygz_search_by_bow_batch_time', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int, f32p]
    sigs.setdefault('ygz_sia_create', {})['argtypes'] = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    sigs.setdefault('ygz_sia_destroy', {})['argtypes'] = [C.c_void_p]
    sigs.setdefault('ygz_sia_run', {})['argtypes'] = [C.c_void_p, C.POINTER(SiaProblem), C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, f32p, f32p, f32p, i32p]
    sigs.setdefault('ygz_sia_run_orb', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SiaProblem), i32p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, f32p, f32p, f32p, i32p]
    sigs.setdefault('ygz_sia_time', {})['argtypes'] = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, f32p]
    sigs.setdefault('ygz_orb_extract_async', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    sigs.setdefault('ygz_orb_wait', {})['argtypes'] = [C.c_void_p]
    sigs.setdefault('ygz_search_by_bow_batch_async', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_float, C.c_int, i32p, i32p, C.c_int, C.c_void_p, C.c_void_p]
    sigs.setdefault('ygz_matcher_set_vocabulary', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int]
    sigs.setdefault('ygz_orb_max_batch', {})['argtypes'] = [C.c_void_p]
    sigs.setdefault('ygz_debug_cos_sin', {})['argtypes'] = [C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p]
    sigs.setdefault('ygz_frontend_time2', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, f32p]
    sigs.setdefault('ygz_frontend_time', {})['argtypes'] = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, f32p]
    sigs.setdefault('ygz_sia_last_trace', {})['argtypes'] = [C.c_void_p, C.c_int, f32p, C.c_int, i32p]
    for name, sig in sigs.items():
        if hasattr(L, name):          # tests/test_abi.py asserts that nothing is missing
            for k, v in sig.items():
                setattr(getattr(L, name), k, v)
    _lib = L
    return L


def missing_symbols():
    L = lib()
    return [n for n in ABI_SYMBOLS if not hasattr(L, n)]


def _check(rc, what):
    if rc < 0:
        raise YgzError("%s failed (%d): %s" % (what, rc, lib().ygz_last_error().decode()))
    return rc


def device_count():
    return lib().ygz_device_count()


def device_info(dev=0):
    name = C.create_string_buffer(256)
    sm = C.c_int()
    l2 = C.c_size_t()
    _check(lib().ygz_device_info(dev, name, C.byref(sm), C.byref(l2)), "ygz_device_info")
    return dict(name=name.value.decode(), sm_count=sm.value, l2_bytes=l2.value)


class ORBextractor:
    """Mirror of ygz::ORBextractor (include/ORBextractor.h:45-172) over the C ABI.

    __call__(image) == operator()(image, mask, keypoints, descriptors)  (src/ORBextractor.cc:970-1028)
    extract_frame(image, existing) == operator()(Frame*, ..., ORBSLAM_KEYPOINT) (:1031-1127)
    """

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7,
                 width=752, height=480, max_batch=1, device=0):
        self.L = lib()
        self.cfg = OrbConfig(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, width, height, max_batch, device)
        self.h = C.c_void_p()
        _check(self.L.ygz_orb_create(C.byref(self.cfg), C.byref(self.h)), "ygz_orb_create")
        self.nlevels, self.width, self.height, self.max_batch = nlevels, width, height, max_batch

    def close(self):
        if getattr(self, "h", None):
            self.L.ygz_orb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- getters of include/ORBextractor.h:82-104
    def GetLevels(self):
        return self.nlevels

    def tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        fpl = np.zeros(n, np.int32)
        _check(self.L.ygz_orb_tables(self.h, _ptr(sc, f32p), _ptr(isc, f32p), _ptr(s2, f32p), _ptr(is2, f32p),
                                     _ptr(fpl, i32p)), "ygz_orb_tables")
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, features_per_level=fpl)

    def GetScaleFactors(self):
        return self.tables()["scale"]

    def GetInverseScaleFactors(self):
        return self.tables()["inv_scale"]

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        _check(self.L.ygz_orb_level_size(self.h, level, C.byref(w), C.byref(h)), "ygz_orb_level_size")
        return w.value, h.value

    def _imgs(self, images):
        images = np.asarray(images)
        if images.ndim == 2:
            images = images[None]
        assert images.dtype == np.uint8 and images.shape[1:] == (self.height, self.width), images.shape
        return np.ascontiguousarray(images)

    def extract_batch(self, images, cap=None):
        """n frames -> list of (keypoints[KP_DTYPE], descriptors[n,32]) ; host in, host out (copies inside)."""
        images = self._imgs(images)
        n = images.shape[0]
        cap = cap or self.cap
        kps = np.zeros((n, cap), KP_DTYPE)
        desc = np.zeros((n, cap, 32), np.uint8)
        counts = np.zeros(n, np.int32)
        _check(self.L.ygz_orb_extract(self.h, images.ctypes.data, n, images.strides[0], images.strides[1],
                                      kps.ctypes.data, desc.ctypes.data, counts.ctypes.data, cap), "ygz_orb_extract")
        return [(kps[f, :counts[f]].copy(), desc[f, :counts[f]].copy()) for f in range(n)]

    def __call__(self, image, mask=None):
        return self.extract_batch(image)[0]

    @property
    def cap(self):
        """Rows per frame of the device output arrays (ygz_orb_max_keypoints); extract_frame / extract_frame_dso may grow it."""
        return self.L.ygz_orb_max_keypoints(self.h)

    def extract_frame(self, image, existing):
        image = self._imgs(image)[0]
        ne = len(existing)
        cap = self.cap + ne
        kps = np.zeros(cap, KP_DTYPE)
        kps[:ne] = existing
        desc = np.zeros((cap, 32), np.uint8)
        cnt = C.c_int()
        _check(self.L.ygz_orb_extract_frame(self.h, image.ctypes.data, image.strides[0], ne, kps.ctypes.data,
                                            desc.ctypes.data, C.byref(cnt), cap), "ygz_orb_extract_frame")
        return kps[:cnt.value].copy(), desc[:cnt.value].copy()

    def extract_frame_dso(self, image, existing, grid_size=-1, cap=None):
        """operator()(Frame*, keys, desc, DSO_KEYPOINT): returns (keypoints incl. the existing ones with refreshed angles,
        descriptors, mnGridSize after the call).  grid_size = mnGridSize before the call (-1: first frame)."""
        image = self._imgs(image)[0]
        ne = len(existing)
        cap = cap or (3 * (image.shape[0] // 7) * (image.shape[1] // 7) + ne)
        kps = np.zeros(cap, KP_DTYPE)
        kps[:ne] = existing
        desc = np.zeros((cap, 32), np.uint8)
        cnt = C.c_int()
        g = C.c_int(int(grid_size))
        _check(self.L.ygz_orb_extract_frame_dso(self.h, image.ctypes.data, image.strides[0], ne, kps.ctypes.data,
                                                desc.ctypes.data, C.byref(cnt), cap, C.byref(g)), "ygz_orb_extract_frame_dso")
        return kps[:cnt.value].copy(), desc[:cnt.value].copy(), g.value

    def extract_frame_dso_batch(self, images, existing, grid_sizes, cap=None):
        """ygz_orb_extract_frame_dso_batch: frame f of B independent streams with its own existing keys and mnGridSize.
        Returns ([(keypoints, descriptors)] per frame, grid sizes after the call)."""
        images = self._imgs(images)
        B = images.shape[0]
        ne = np.array([len(e) for e in existing], np.int32)
        full = 3 * (images.shape[1] // 7) * (images.shape[2] // 7) + int(ne.max(initial=0))     # three winners in every 7 x 7 cell
        # the selection stops near nfeatures new keys (overshoot below 3 x): try a tight buffer first, the worst case on YGZ_ERR_CAPACITY
        for c in ([cap] if cap else [min(full, 4 * self.cfg.nfeatures + int(ne.max(initial=0))), full]):
            kps = np.zeros((B, c), KP_DTYPE)
            for f in range(B):
                kps[f, :ne[f]] = existing[f]
            desc = np.zeros((B, c, 32), np.uint8)
            counts = np.zeros(B, np.int32)
            g = np.ascontiguousarray(grid_sizes, np.int32).copy()
            rc = self.L.ygz_orb_extract_frame_dso_batch(self.h, images.ctypes.data, B, images.strides[0], images.strides[1], _ptr(ne, i32p),
                                                        kps.ctypes.data, desc.ctypes.data, _ptr(counts, i32p), c, _ptr(g, i32p))
            if rc != -3 or c == full or cap:            # YGZ_ERR_CAPACITY
                break
        _check(rc, "ygz_orb_extract_frame_dso_batch")
        return [(kps[f, :counts[f]].copy(), desc[f, :counts[f]].copy()) for f in range(B)], g

    def ComputePyramid(self, images):
        images = self._imgs(images)
        _check(self.L.ygz_orb_compute_pyramid(self.h, images.ctypes.data, images.shape[0], images.strides[0],
                                              images.strides[1]), "ygz_orb_compute_pyramid")

    def level(self, frame, level, border=0):
        """mvImagePyramid[level] of frame `frame` (continuous copy)."""
        w, h = self.level_size(level)
        out = np.zeros((h + 2 * border, w + 2 * border), np.uint8)
        _check(self.L.ygz_orb_get_level(self.h, frame, level, border, out.ctypes.data), "ygz_orb_get_level")
        return out

    def blurred_level(self, frame, level):
        """GaussianBlur(mvImagePyramid[level]) as the last extraction left it on the device (parity instrumentation)."""
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        _check(self.L.ygz_orb_get_blurred_level(self.h, frame, level, out.ctypes.data), "ygz_orb_get_blurred_level")
        return out

    def candidates(self, frame, level, cap=400000):
        out = np.zeros(cap, KP_DTYPE)
        n = self.L.ygz_orb_get_candidates(self.h, frame, level, out.ctypes.data, cap)
        _check(n, "ygz_orb_get_candidates")
        return out[:n].copy()

    # -- split phases (device-resident timing)
    def upload(self, images):
        images = self._imgs(images)
        _check(self.L.ygz_orb_upload(self.h, images.ctypes.data, images.shape[0], images.strides[0],
                                     images.strides[1]), "ygz_orb_upload")
        return images.shape[0]

    def upload_ptr(self, ptr, n):
        _check(self.L.ygz_orb_upload(self.h, ptr, n, self.width * self.height, self.width), "ygz_orb_upload")

    def run(self, n):
        _check(self.L.ygz_orb_run(self.h, n), "ygz_orb_run")

    def sync(self):
        _check(self.L.ygz_orb_sync(self.h), "ygz_orb_sync")

    def download(self, n, cap=None):
        cap = cap or self.cap
        kps = np.zeros((n, cap), KP_DTYPE)
        desc = np.zeros((n, cap, 32), np.uint8)
        counts = np.zeros(n, np.int32)
        _check(self.L.ygz_orb_download(self.h, n, kps.ctypes.data, desc.ctypes.data, counts.ctypes.data, cap),
               "ygz_orb_download")
        return kps, desc, counts

    def download_ptr(self, n, kps_ptr, desc_ptr, counts_ptr, cap):
        _check(self.L.ygz_orb_download(self.h, n, kps_ptr, desc_ptr, counts_ptr, cap), "ygz_orb_download")

    def time(self, n, iters):
        ms = np.zeros(6, np.float32)
        launches = C.c_int()
        _check(self.L.ygz_orb_time(self.h, n, iters, _ptr(ms, f32p), C.byref(launches)), "ygz_orb_time")
        return dict(pyramid=float(ms[0]), fast_cells=float(ms[1]), octree=float(ms[2]), blur=float(ms[3]),
                    describe=float(ms[4]), total=float(ms[5]), launches=launches.value)

    def counters(self, reset=False):
        c = Counters()
        _check(self.L.ygz_orb_counters(self.h, C.byref(c), int(reset)), "ygz_orb_counters")
        return {k: int(getattr(c, k)) for k, _ in Counters._fields_}


# ---------------------------------------------------------------------------------------------- views
class _Keep:
    """ctypes struct + the numpy arrays it points into (kept alive together)."""

    def __init__(self, struct, refs):
        self.struct = struct
        self.refs = refs

    def ptr(self, cls=None):
        if cls is None:
            return C.byref(self.struct)
        return C.cast(C.byref(self.struct), C.POINTER(cls))


def make_frame_view(keys, desc, map_point=None, u_right=None, outlier=None, K=(458.654, 457.296, 367.215, 248.375),
                    mb=0.0, mbf=0.0, bounds=(0.0, 0.0, 752.0, 480.0), scale_factors=None, Rcw=None, tcw=None,
                    node_of_feature=None, cls=None):
    """Flat view of a Frame / KeyFrame (include/ygz_b200.h ygz_frame_view).  node_of_feature (uint32 per feature,
    0xFFFFFFFF = none) is turned into the CSR form of a DBoW2::FeatureVector (nodes ascending, indices ascending)."""
    cls = cls or FrameView
    keys = np.ascontiguousarray(keys, KP_DTYPE)
    n = len(keys)
    desc = np.ascontiguousarray(desc, np.uint8).reshape(n, 32)
    mp = np.full(n, -1, np.int32) if map_point is None else np.ascontiguousarray(map_point, np.int32)
    sf = np.ascontiguousarray(scale_factors if scale_factors is not None else 1.2 ** np.arange(8), np.float32)
    isf = (np.float32(1.0) / sf).astype(np.float32)
    v = cls()
    refs = [keys, desc, mp, sf, isf]
    v.n = n
    v.keys = keys.ctypes.data
    v.desc = _ptr(desc, u8p)
    v.map_point = _ptr(mp, i32p)
    if u_right is not None:
        u_right = np.ascontiguousarray(u_right, np.float32); refs.append(u_right); v.u_right = _ptr(u_right, f32p)
    if outlier is not None:
        outlier = np.ascontiguousarray(outlier, np.uint8); refs.append(outlier); v.outlier = _ptr(outlier, u8p)
    v.fx, v.fy, v.cx, v.cy = K
    v.mb, v.mbf = mb, mbf
    v.min_x, v.min_y, v.max_x, v.max_y = bounds
    v.n_levels = len(sf)
    v.scale_factors = _ptr(sf, f32p)
    v.inv_scale_factors = _ptr(isf, f32p)
    v.log_scale_factor = float(np.log(np.float32(sf[1]))) if len(sf) > 1 else 0.0
    R = np.eye(3, dtype=np.float32) if Rcw is None else np.asarray(Rcw, np.float32)
    t = np.zeros(3, np.float32) if tcw is None else np.asarray(tcw, np.float32)
    v.Rcw = (C.c_float * 9)(*R.reshape(-1))
    v.tcw = (C.c_float * 3)(*t)
    if node_of_feature is not None:
        nof = np.asarray(node_of_feature, np.uint32)[:n]
        valid = np.nonzero(nof != 0xFFFFFFFF)[0]
        order = valid[np.lexsort((valid, nof[valid]))]
        nodes, counts = np.unique(nof[order], return_counts=True)
        node_id = np.ascontiguousarray(nodes, np.uint32)
        offset = np.zeros(len(nodes) + 1, np.int32)
        offset[1:] = np.cumsum(counts)
        index = np.ascontiguousarray(order, np.uint32)
        refs += [node_id, offset, index]
        v.fv_nodes = len(nodes)
        v.fv_node_id = _ptr(node_id, u32p)
        v.fv_offset = _ptr(offset, i32p)
        v.fv_index = _ptr(index, u32p)
    k = _Keep(v, refs)
    k.map_point = mp
    return k


def make_map_points(world_pos, desc, bad=None, observations=None, min_dist=None, max_dist=None, track=None, cls=None, normal=None):
    cls = cls or MapPoints
    m = len(world_pos)
    v = cls()
    wp = np.ascontiguousarray(world_pos, np.float32).reshape(m, 3)
    d = np.ascontiguousarray(desc, np.uint8).reshape(m, 32)
    refs = [wp, d]
    v.m = m
    v.world_pos = _ptr(wp, f32p)
    v.desc = _ptr(d, u8p)

    def opt(name, arr, dt, pt):
        if arr is not None:
            a = np.ascontiguousarray(arr, dt)
            refs.append(a)
            setattr(v, name, _ptr(a, pt))
    opt("bad", bad, np.uint8, u8p)
    opt("observations", observations, np.int32, i32p)
    opt("min_dist", min_dist if min_dist is not None else np.full(m, 0.1), np.float32, f32p)
    opt("max_dist", max_dist if max_dist is not None else np.full(m, 100.0), np.float32, f32p)
    track = track or {}
    opt("track_in_view", track.get("in_view", np.ones(m)), np.uint8, u8p)
    opt("track_proj_x", track.get("x", np.zeros(m)), np.float32, f32p)
    opt("track_proj_y", track.get("y", np.zeros(m)), np.float32, f32p)
    opt("track_proj_xr", track.get("xr", np.zeros(m)), np.float32, f32p)
    opt("track_scale_level", track.get("level", np.zeros(m)), np.int32, i32p)
    opt("track_view_cos", track.get("view_cos", np.ones(m)), np.float32, f32p)
    opt("normal", normal, np.float32, f32p)
    return _Keep(v, refs)


class ORBmatcher:
    """Mirror of ygz::ORBmatcher (include/ORBmatcher.h:38-220): ORBmatcher(nnratio=0.6, checkOri=True)."""
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30

    def __init__(self, nnratio=0.6, checkOri=True, device=0, max_queries=8192, max_targets=8192, max_pairs=1 << 20):
        self.L = lib()
        self.nnratio, self.checkOri = float(nnratio), bool(checkOri)
        self.h = C.c_void_p()
        _check(self.L.ygz_matcher_create(device, max_queries, max_targets, max_pairs, C.byref(self.h)), "ygz_matcher_create")

    def close(self):
        if getattr(self, "h", None):
            self.L.ygz_matcher_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_vocabulary(self, vocab, levelsup=4):
        """Frame::ComputeBoW on the device: the batched searches take their FeatureVector node ids from `vocab`
        (an ORBVocabulary on the same device) instead of node_of_feature / the synthetic stand-in.  None detaches."""
        self._vocab = vocab                     # keep the handle alive
        _check(self.L.ygz_matcher_set_vocabulary(self.h, vocab.h if vocab is not None else None, levelsup),
               "ygz_matcher_set_vocabulary")

    @staticmethod
    def DescriptorDistance(a, b):
        a = np.frombuffer(bytes(a), np.uint8) if isinstance(a, (bytes, bytearray)) else np.ascontiguousarray(a, np.uint8)
        b = np.frombuffer(bytes(b), np.uint8) if isinstance(b, (bytes, bytearray)) else np.ascontiguousarray(b, np.uint8)
        return lib().ygz_descriptor_distance(a.ctypes.data, b.ctypes.data)

    def bf_top2(self, q, t):
        q = np.ascontiguousarray(q, np.uint8)
        t = np.ascontiguousarray(t, np.uint8)
        bi, bd, sd = (np.zeros(len(q), np.int32) for _ in range(3))
        _check(self.L.ygz_hamming_bf_top2(self.h, q.ctypes.data, len(q), t.ctypes.data, len(t), _ptr(bi, i32p),
                                          _ptr(bd, i32p), _ptr(sd, i32p)), "ygz_hamming_bf_top2")
        return bi, bd, sd

    def pairs(self, q, t, pq, pt):
        q = np.ascontiguousarray(q, np.uint8)
        t = np.ascontiguousarray(t, np.uint8)
        pq = np.ascontiguousarray(pq, np.int32)
        pt = np.ascontiguousarray(pt, np.int32)
        d = np.zeros(len(pq), np.int32)
        _check(self.L.ygz_hamming_pairs(self.h, q.ctypes.data, len(q), t.ctypes.data, len(t), _ptr(pq, i32p),
                                        _ptr(pt, i32p), len(pq), _ptr(d, i32p)), "ygz_hamming_pairs")
        return d

    def SearchByBoW(self, kf, mps, f):
        out = np.zeros(max(f.struct.n, 1), np.int32)
        n = _check(self.L.ygz_search_by_bow(self.h, self.nnratio, int(self.checkOri), kf.ptr(), mps.ptr(), f.ptr(),
                                            _ptr(out, i32p)), "ygz_search_by_bow")
        return n, out[:f.struct.n]

    def SearchByProjectionPoints(self, f, mps, query_ids, th, checkLevel=True):
        q = np.ascontiguousarray(query_ids, np.int32)
        return _check(self.L.ygz_search_by_projection_points(self.h, self.nnratio, f.ptr(), mps.ptr(), _ptr(q, i32p),
                                                             len(q), th, int(checkLevel)), "ygz_search_by_projection_points")

    def SearchByProjectionLast(self, cur, last, mps, th, mono=True, checkLevel=True):
        return _check(self.L.ygz_search_by_projection_last(self.h, int(self.checkOri), cur.ptr(), last.ptr(), mps.ptr(),
                                                           th, int(mono), int(checkLevel)), "ygz_search_by_projection_last")

    def SearchByProjectionKF(self, cur, kf, mps, already_found, th, orb_dist):
        af = None if already_found is None else np.ascontiguousarray(already_found, np.uint8)
        return _check(self.L.ygz_search_by_projection_kf(self.h, int(self.checkOri), cur.ptr(), kf.ptr(), mps.ptr(),
                                                         _ptr(af, u8p), th, orb_dist), "ygz_search_by_projection_kf")

    def SearchForInitialization(self, f1, f2, prev_matched, window=100):
        prev = np.ascontiguousarray(prev_matched, np.float32).copy()
        m12 = np.zeros(max(f1.struct.n, 1), np.int32)
        n = _check(self.L.ygz_search_for_initialization(self.h, self.nnratio, int(self.checkOri), f1.ptr(), f2.ptr(),
                                                        _ptr(prev, f32p), _ptr(m12, i32p), window),
                   "ygz_search_for_initialization")
        return n, m12[:f1.struct.n], prev

    def SearchByBoWKF(self, kf1, kf2, mps):
        """SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) :480-595.  Returns (nmatches, matches12[kf1.n])."""
        out = np.zeros(max(kf1.struct.n, 1), np.int32)
        n = self.L.ygz_search_by_bow_kf(self.h, self.nnratio, int(self.checkOri), kf1.ptr(), kf2.ptr(), mps.ptr(), _ptr(out, i32p))
        if n < 0:
            _check(n, "ygz_search_by_bow_kf")
        return n, out[:kf1.struct.n]

    def SearchForTriangulation(self, kf1, kf2, F12, epipole, level_sigma2, only_stereo=False):
        """SearchForTriangulation :597-746.  Returns (nmatches, pairs[n, 2])."""
        F12 = np.ascontiguousarray(F12, np.float32).reshape(9)
        ls2 = np.ascontiguousarray(level_sigma2, np.float32)
        pairs = np.zeros((max(kf1.struct.n, 1), 2), np.int32)
        n = self.L.ygz_search_for_triangulation(self.h, int(self.checkOri), kf1.ptr(), kf2.ptr(), _ptr(F12, f32p), float(epipole[0]),
                                                float(epipole[1]), _ptr(ls2, f32p), int(only_stereo), _ptr(pairs, i32p))
        if n < 0:
            _check(n, "ygz_search_for_triangulation")
        return n, pairs[:n].copy()

    def SearchByBoWBatch(self, extractor, kf_idx, f_idx, node_of_feature):
        """Device-resident SearchByBoW over frame pairs of an extractor batch (outputs already in HBM)."""
        kf_idx = np.ascontiguousarray(kf_idx, np.int32)
        f_idx = np.ascontiguousarray(f_idx, np.int32)
        cap = extractor.cap
        nof = None
        if node_of_feature is not None:
            nof = np.ascontiguousarray(node_of_feature, np.uint32)
            assert nof.ndim == 2 and nof.shape[1] == cap
        out = np.zeros((len(kf_idx), cap), np.int32)
        nm = np.zeros(len(kf_idx), np.int32)
        _check(self.L.ygz_search_by_bow_batch(self.h, extractor.h, self.nnratio, int(self.checkOri), _ptr(kf_idx, i32p),
                                              _ptr(f_idx, i32p), len(kf_idx), _ptr(nof, u32p), _ptr(out, i32p),
                                              _ptr(nm, i32p)), "ygz_search_by_bow_batch")
        return nm, out

    def time_frontend2(self, extractor, other_matcher, other_extractor, n_frames, iters):
        ms = C.c_float()
        _check(self.L.ygz_frontend_time2(extractor.h, self.h, other_extractor.h, other_matcher.h, n_frames, iters,
                                         self.nnratio, int(self.checkOri), C.byref(ms)), "ygz_frontend_time2")
        return ms.value

    @staticmethod
    def time_frontend_n(pairs, n_frames, iters, nnratio, check_ori):
        """pairs: [(ORBextractor, ORBmatcher), ...] — k handle pairs in flight on k streams; returns total ms over iters steps."""
        k = len(pairs)
        orbs = (C.c_void_p * k)(*[p[0].h for p in pairs]); hs = (C.c_void_p * k)(*[p[1].h for p in pairs])
        ms = C.c_float()
        L = lib()
        L.ygz_frontend_time_n.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, f32p]
        _check(L.ygz_frontend_time_n(orbs, hs, k, n_frames, iters, float(nnratio), int(check_ori), C.byref(ms)), "ygz_frontend_time_n")
        return ms.value

    def time_frontend(self, extractor, n_frames, iters):
        ms = C.c_float()
        _check(self.L.ygz_frontend_time(extractor.h, self.h, n_frames, iters, self.nnratio, int(self.checkOri),
                                        C.byref(ms)), "ygz_frontend_time")
        return ms.value

    def SearchByProjectionLastBatch(self, ex, last_frame, cur_frame, world_pos, valid, pose, K=(458.654, 457.296, 367.215, 248.375),
                                    bounds=(0.0, 0.0, 752.0, 480.0), mb=0.0, mbf=0.0, mp_desc=None, u_right=None, th=15.0, mono=True,
                                    check_level=True, check_ori=True):
        """Batched device-resident SearchByProjection(Current, Last) (src/ORBmatcher.cc:1218-1350) over frame pairs of the extractor
        batch.  world_pos [B, cap, 3], valid [B, cap], pose [B, 12].  Returns (nmatches[n], matches[n, cap], rounds[n])."""
        lf = np.ascontiguousarray(last_frame, np.int32); cf = np.ascontiguousarray(cur_frame, np.int32)
        n, cap = len(lf), ex.cap
        cam = ProjectionCamera(*K, mb, mbf, *bounds)
        wp = np.ascontiguousarray(world_pos, np.float32); va = np.ascontiguousarray(valid, np.uint8); po = np.ascontiguousarray(pose, np.float32)
        md = None if mp_desc is None else np.ascontiguousarray(mp_desc, np.uint8)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        out = np.zeros((n, cap), np.int32); nm = np.zeros(n, np.int32); rd = np.zeros(n, np.int32)
        _check(self.L.ygz_search_by_projection_last_batch(self.h, ex.h, C.byref(cam), n, _ptr(lf, i32p), _ptr(cf, i32p), _ptr(wp, f32p), _ptr(va, u8p),
                                                          _ptr(po, f32p), _ptr(md, u8p), _ptr(ur, f32p), float(th), int(mono), int(check_level),
                                                          int(check_ori), _ptr(out, i32p), _ptr(nm, i32p), _ptr(rd, i32p)),
               "ygz_search_by_projection_last_batch")
        return nm, out, rd

    def SearchByProjectionPointsBatch(self, ex, frame, proj_x, proj_y, proj_xr, view_cos, level, usable, mp_desc, occupied=None, u_right=None,
                                      K=(458.654, 457.296, 367.215, 248.375), bounds=(0.0, 0.0, 752.0, 480.0), mb=0.0, mbf=0.0, nnratio=0.8,
                                      th=1.0, check_level=True):
        """Batched device-resident SearchByProjection(F, vpMapPoints) (src/ORBmatcher.cc:43-126): arrays [n, M] per pair.
        Returns (nmatches[n], matches[n, cap], rounds[n])."""
        fr = np.ascontiguousarray(frame, np.int32)
        n, cap = len(fr), ex.cap
        px = np.ascontiguousarray(proj_x, np.float32); M = px.shape[1]
        py = np.ascontiguousarray(proj_y, np.float32); pxr = np.ascontiguousarray(proj_xr, np.float32)
        vc = np.ascontiguousarray(view_cos, np.float32); lv = np.ascontiguousarray(level, np.int32); us = np.ascontiguousarray(usable, np.uint8)
        md = np.ascontiguousarray(mp_desc, np.uint8)
        oc = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        cam = ProjectionCamera(*K, mb, mbf, *bounds)
        out = np.zeros((n, cap), np.int32); nm = np.zeros(n, np.int32); rd = np.zeros(n, np.int32)
        _check(self.L.ygz_search_by_projection_points_batch(self.h, ex.h, C.byref(cam), n, _ptr(fr, i32p), M, _ptr(px, f32p), _ptr(py, f32p),
                                                            _ptr(pxr, f32p), _ptr(vc, f32p), _ptr(lv, i32p), _ptr(us, u8p), _ptr(md, u8p),
                                                            _ptr(oc, u8p), _ptr(ur, f32p), float(nnratio), float(th), int(check_level),
                                                            _ptr(out, i32p), _ptr(nm, i32p), _ptr(rd, i32p)),
               "ygz_search_by_projection_points_batch")
        return nm, out, rd

    def time_projection_batch(self, ex, iters):
        ms = C.c_float()
        _check(self.L.ygz_search_by_projection_batch_time(self.h, ex.h, iters, C.byref(ms)), "ygz_search_by_projection_batch_time")
        return ms.value

    def time_bow_batch(self, extractor, n_pairs, iters):
        ms = C.c_float()
        _check(self.L.ygz_search_by_bow_batch_time(self.h, extractor.h, self.nnratio, int(self.checkOri), n_pairs, iters,
                                                   C.byref(ms)), "ygz_search_by_bow_batch_time")
        return ms.value


def make_sia_problem(ref_levels, cur_levels, inv_scale, K, ref_px, world_xyz, valid=None, ref_pose=None, cur_pose=None,
                     cls=None):
    """ygz_sia_problem from numpy: ref/cur level images (lists of 2-D uint8), poses as (q[4] xyzw, t[3])."""
    cls = cls or SiaProblem
    nl = len(ref_levels)
    ref_levels = [np.ascontiguousarray(a, np.uint8) for a in ref_levels]
    cur_levels = [np.ascontiguousarray(a, np.uint8) for a in cur_levels]
    lw = np.array([a.shape[1] for a in ref_levels], np.int32)
    lh = np.array([a.shape[0] for a in ref_levels], np.int32)
    isf = np.ascontiguousarray(inv_scale, np.float32)
    px = np.ascontiguousarray(ref_px, np.float32).reshape(-1, 2)
    xyz = np.ascontiguousarray(world_xyz, np.float32).reshape(-1, 3)
    n = len(px)
    val = np.ones(n, np.uint8) if valid is None else np.ascontiguousarray(valid, np.uint8)
    refp = (u8p * nl)(*[_ptr(a, u8p) for a in ref_levels])
    curp = (u8p * nl)(*[_ptr(a, u8p) for a in cur_levels])
    v = cls()
    v.n_levels = nl
    v.ref_levels = C.cast(refp, C.POINTER(u8p))
    v.cur_levels = C.cast(curp, C.POINTER(u8p))
    v.level_w = _ptr(lw, i32p)
    v.level_h = _ptr(lh, i32p)
    v.inv_scale_factors = _ptr(isf, f32p)
    v.fx, v.fy, v.cx, v.cy = K
    v.n = n
    v.ref_px = _ptr(px, f32p)
    v.world_xyz = _ptr(xyz, f32p)
    v.valid = _ptr(val, u8p)
    rq, rt = ref_pose if ref_pose is not None else ((0, 0, 0, 1), (0, 0, 0))
    cq, ct = cur_pose if cur_pose is not None else ((0, 0, 0, 1), (0, 0, 0))
    v.ref_q = (C.c_float * 4)(*rq)
    v.ref_t = (C.c_float * 3)(*rt)
    v.cur_q = (C.c_float * 4)(*cq)
    v.cur_t = (C.c_float * 3)(*ct)
    return _Keep(v, [ref_levels, cur_levels, lw, lh, isf, px, xyz, val, refp, curp])


def make_sia_features(K, ref_px, world_xyz, valid=None, ref_pose=None, cur_pose=None):
    """ygz_sia_problem carrying only what ygz_sia_run_orb reads: camera, features, poses (the images stay in the extractor handles)."""
    px = np.ascontiguousarray(ref_px, np.float32).reshape(-1, 2)
    xyz = np.ascontiguousarray(world_xyz, np.float32).reshape(-1, 3)
    n = len(px)
    val = np.ones(n, np.uint8) if valid is None else np.ascontiguousarray(valid, np.uint8)
    v = SiaProblem()
    v.n_levels = 0
    v.fx, v.fy, v.cx, v.cy = K
    v.n = n
    v.ref_px = _ptr(px, f32p)
    v.world_xyz = _ptr(xyz, f32p)
    v.valid = _ptr(val, u8p)
    rq, rt = ref_pose if ref_pose is not None else ((0, 0, 0, 1), (0, 0, 0))
    cq, ct = cur_pose if cur_pose is not None else ((0, 0, 0, 1), (0, 0, 0))
    v.ref_q = (C.c_float * 4)(*rq)
    v.ref_t = (C.c_float * 3)(*rt)
    v.cur_q = (C.c_float * 4)(*cq)
    v.cur_t = (C.c_float * 3)(*ct)
    return _Keep(v, [px, xyz, val])


class ORBVocabulary:
    """Mirror of ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (include/ORBVocabulary.h) for the calls
    the front end makes: loadFromBinaryFile and transform(features, BowVector, FeatureVector, levelsup)."""

    def __init__(self, path=None, data=None, device=0):
        self.L_ = lib()
        self.h = C.c_void_p()
        if path is not None:
            _check(self.L_.ygz_vocab_load(str(path).encode(), device, C.byref(self.h)), "ygz_vocab_load")
        else:
            data = np.ascontiguousarray(np.frombuffer(data, np.uint8))
            _check(self.L_.ygz_vocab_from_memory(data.ctypes.data, data.size, device, C.byref(self.h)), "ygz_vocab_from_memory")
        v = [C.c_int() for _ in range(6)]
        _check(self.L_.ygz_vocab_info(self.h, *[C.byref(x) for x in v]), "ygz_vocab_info")
        self.k, self.L, self.scoring, self.weighting, self.n_nodes, self.n_words = [x.value for x in v]

    def close(self):
        if getattr(self, "h", None):
            self.L_.ygz_vocab_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transform_features(self, desc, levelsup=4):
        """per descriptor: (word_id, weight, node_id at level L - levelsup)"""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        wid = np.zeros(n, np.int32); w = np.zeros(n, np.float32); nid = np.zeros(n, np.int32)
        _check(self.L_.ygz_vocab_transform(self.h, desc.ctypes.data, n, levelsup, _ptr(wid, i32p), _ptr(w, f32p), _ptr(nid, i32p)),
               "ygz_vocab_transform")
        return wid, w, nid

    def transform(self, desc, levelsup=4):
        """transform(features, BowVector&, FeatureVector&, levelsup) (TemplatedVocabulary.h:1151-1227, TF-IDF + L1):
        returns (bow_ids ascending, bow_vals float64 L1-normalised, fv: dict node -> feature indices ascending)."""
        wid, w, nid = self.transform_features(desc, levelsup)
        keep = np.nonzero(w > 0)[0]                       # w > 0: not a stopped word
        ids, inv = np.unique(wid[keep], return_inverse=True)
        vals = np.zeros(len(ids), np.float64)
        np.add.at(vals, inv, w[keep].astype(np.float64))  # addWeight in feature order (sequential double sums)
        norm = float(np.cumsum(np.abs(vals))[-1]) if len(vals) else 0.0      # map iteration order, sequential
        if norm > 0.0:
            vals = vals / norm
        fv = {}
        for i in keep:
            fv.setdefault(int(nid[i]), []).append(int(i))
        return ids.astype(np.int32), vals, dict(sorted(fv.items()))

    def transform_orb(self, ex, n_frames, levelsup=4, want=("node",), timed=False):
        cap = ex.cap
        wid = np.zeros((n_frames, cap), np.int32) if "word" in want else None
        w = np.zeros((n_frames, cap), np.float32) if "weight" in want else None
        nid = np.zeros((n_frames, cap), np.int32) if "node" in want else None
        ms = C.c_float()
        _check(self.L_.ygz_vocab_transform_orb(self.h, ex.h, n_frames, levelsup, _ptr(wid, i32p), _ptr(w, f32p), _ptr(nid, i32p),
                                               C.byref(ms) if timed else None), "ygz_vocab_transform_orb")
        return dict(word=wid, weight=w, node=nid, ms=ms.value)


def ComputeStereoMatches(ex_left, ex_right, keysL, descL, keysR, descR, mb, mbf, frame=0):
    """Frame::ComputeStereoMatches (src/Frame.cc:509-682) on the pyramids held by two ORBextractor handles.
    Returns (mvuRight[N], mvDepth[N], n_matches)."""
    keysL = np.ascontiguousarray(keysL, KP_DTYPE); keysR = np.ascontiguousarray(keysR, KP_DTYPE)
    descL = np.ascontiguousarray(descL, np.uint8); descR = np.ascontiguousarray(descR, np.uint8)
    u = np.zeros(len(keysL), np.float32); z = np.zeros(len(keysL), np.float32)
    n = lib().ygz_compute_stereo_matches(ex_left.h, ex_right.h, frame, keysL.ctypes.data, descL.ctypes.data, len(keysL),
                                         keysR.ctypes.data, descR.ctypes.data, len(keysR), float(mb), float(mbf),
                                         _ptr(u, f32p), _ptr(z, f32p))
    if n < 0:
        _check(n, "ygz_compute_stereo_matches")
    return u, z, n


def StereoBatch(ex_left, ex_right, n_frames, mb, mbf, brute_force=False):
    """Batched, device-resident Frame::ComputeStereoMatches over the n_frames stereo pairs the two extractor handles hold (their last
    batch), median cut included, plus (brute_force) the dense best / second-best right descriptor of every left key (BASELINE C5).
    Returns dict(u_right[n, cap], depth[n, cap], n_matches[n], bf_best / bf_dist / bf_second [n, cap] or None)."""
    cap = ex_left.cap
    u = np.zeros((n_frames, cap), np.float32); z = np.zeros((n_frames, cap), np.float32); nm = np.zeros(n_frames, np.int32)
    bf = [np.zeros((n_frames, cap), np.int32) for _ in range(3)] if brute_force else [None] * 3
    _check(lib().ygz_stereo_batch(ex_left.h, ex_right.h, n_frames, float(mb), float(mbf), _ptr(u, f32p), _ptr(z, f32p), _ptr(nm, i32p),
                                  _ptr(bf[0], i32p), _ptr(bf[1], i32p), _ptr(bf[2], i32p)), "ygz_stereo_batch")
    return dict(u_right=u, depth=z, n_matches=nm, bf_best=bf[0], bf_dist=bf[1], bf_second=bf[2])


def time_stereo_frontend(ex_left, ex_right, n_frames, iters, mb, mbf, brute_force=True):
    """Device-timed C5 steps (both extractors + batched stereo association + median cut + brute force), inputs resident.
    Returns (total ms over iters, stage ms of the last iteration {extract, match, median_cut, brute_force})."""
    ms = C.c_float()
    st = np.zeros(4, np.float32)
    _check(lib().ygz_stereo_frontend_time(ex_left.h, ex_right.h, n_frames, iters, float(mb), float(mbf), int(brute_force), C.byref(ms),
                                          _ptr(st, f32p)), "ygz_stereo_frontend_time")
    return ms.value, dict(zip(("extract", "match", "median_cut", "brute_force"), [float(x) for x in st]))


def make_direct_frame(levels, q=(0, 0, 0, 1), t=(0, 0, 0), cls=None):
    """ygz_direct_frame from a list of 2-D uint8 level images and a pose (q xyzw, t)."""
    cls = cls or DirectFrame
    levels = [np.ascontiguousarray(a, np.uint8) for a in levels]
    f = cls()
    for l, a in enumerate(levels):
        f.level[l] = _ptr(a, u8p)
        f.w[l], f.h[l], f.pitch[l] = a.shape[1], a.shape[0], a.strides[0]
    f.q = (C.c_float * 4)(*[float(x) for x in q])
    f.t = (C.c_float * 3)(*[float(x) for x in t])
    return _Keep(f, levels)


class DirectMatcher:
    """Batched ORBmatcher::FindDirectProjection (src/ORBmatcher.cc:1575-1602): keyframe / current-frame pyramids live
    in device slots, one call searches all (map point, keyframe) items of a frame."""

    def __init__(self, K, scale, inv_scale, inv_level_sigma2_1, max_slots=8, max_items=4096,
                 max_pyramid_bytes=4 * 752 * 480, device=0):
        self.L = lib()
        self.K = np.ascontiguousarray(K, np.float32)
        self.scale = np.ascontiguousarray(scale, np.float32)
        self.inv_scale = np.ascontiguousarray(inv_scale, np.float32)
        self.inv_sigma2_1 = float(inv_level_sigma2_1)
        self.nlevels = len(self.scale)
        self.h = C.c_void_p()
        _check(self.L.ygz_direct_create(device, max_slots, max_items, max_pyramid_bytes, C.byref(self.h)), "ygz_direct_create")

    def close(self):
        if getattr(self, "h", None):
            self.L.ygz_direct_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_frame(self, slot, frame):
        _check(self.L.ygz_direct_upload_frame(self.h, slot, C.byref(frame.struct), self.nlevels), "ygz_direct_upload_frame")

    def set_pose(self, slot, q, t):
        q = np.ascontiguousarray(q, np.float32); t = np.ascontiguousarray(t, np.float32)
        _check(self.L.ygz_direct_set_pose(self.h, slot, _ptr(q, f32p), _ptr(t, f32p)), "ygz_direct_set_pose")

    def FindDirectProjection(self, cur_slot, items):
        """items: structured array of DIRECT_ITEM_DTYPE.  Returns (px[n,2], level[n], ok[n])."""
        items = np.ascontiguousarray(items, DIRECT_ITEM_DTYPE)
        n = len(items)
        px = np.zeros((n, 2), np.float32); lvl = np.zeros(n, np.int32); ok = np.zeros(n, np.uint8)
        _check(self.L.ygz_find_direct_projection(self.h, cur_slot, _ptr(self.K, f32p), _ptr(self.scale, f32p),
                                                 _ptr(self.inv_scale, f32p), self.inv_sigma2_1, self.nlevels,
                                                 items.ctypes.data, n, _ptr(px, f32p), _ptr(lvl, i32p), _ptr(ok, u8p)),
               "ygz_find_direct_projection")
        return px, lvl, ok

    def time(self, iters):
        ms = C.c_float()
        _check(self.L.ygz_direct_time(self.h, iters, C.byref(ms)), "ygz_direct_time")
        return ms.value


class SparseImgAlign:
    """Mirror of ygz::SparseImgAlign (include/SparseImageAlign.h:14-113): SparseImgAlign(max_level, min_level, n_iter=10)."""

    def __init__(self, max_level, min_level, n_iter=10, device=0, max_problems=1, max_features=2048,
                 max_level_pixels=752 * 480 * 2):
        self.L = lib()
        self.max_level, self.min_level, self.n_iter = max_level, min_level, n_iter
        self.eps = 0.000001                       # eps_, src/SparseImageAlign.cc:17
        self.h = C.c_void_p()
        _check(self.L.ygz_sia_create(device, max_problems, max_features, max_level_pixels, C.byref(self.h)), "ygz_sia_create")

    def close(self):
        if getattr(self, "h", None):
            self.L.ygz_sia_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_batch(self, problems):
        n = len(problems)
        arr = (SiaProblem * n)(*[p.struct for p in problems])
        q = np.zeros((n, 4), np.float32)
        t = np.zeros((n, 3), np.float32)
        H = np.zeros((n, 36), np.float32)
        nt = np.zeros(n, np.int32)
        _check(self.L.ygz_sia_run(self.h, arr, n, self.max_level, self.min_level, self.n_iter, self.eps, _ptr(q, f32p),
                                  _ptr(t, f32p), _ptr(H, f32p), _ptr(nt, i32p)), "ygz_sia_run")
        return q, t, H, nt

    def run_batch_orb(self, orb_ref, orb_cur, problems, ref_frame, cur_frame):
        """ygz_sia_run_orb: level images read in place from two ORBextractor handles (frames ref_frame[p] / cur_frame[p]); the
        problems come from make_sia_features (or make_sia_problem - their level fields are ignored)."""
        n = len(problems)
        arr = (SiaProblem * n)(*[p.struct for p in problems])
        rf = np.ascontiguousarray(ref_frame, np.int32)
        cf = np.ascontiguousarray(cur_frame, np.int32)
        assert len(rf) == n and len(cf) == n
        q = np.zeros((n, 4), np.float32)
        t = np.zeros((n, 3), np.float32)
        H = np.zeros((n, 36), np.float32)
        nt = np.zeros(n, np.int32)
        _check(self.L.ygz_sia_run_orb(self.h, orb_ref.h, orb_cur.h, arr, _ptr(rf, i32p), _ptr(cf, i32p), n, self.max_level, self.min_level,
                                      self.n_iter, self.eps, _ptr(q, f32p), _ptr(t, f32p), _ptr(H, f32p), _ptr(nt, i32p)), "ygz_sia_run_orb")
        return q, t, H, nt

    def run(self, problem):
        q, t, H, nt = self.run_batch([problem])
        return int(nt[0]), q[0], t[0], H[0].reshape(6, 6)

    def trace(self, problem_index=0, cap=192):
        out = np.zeros((cap, 8), np.float32)
        n = C.c_int()
        _check(self.L.ygz_sia_last_trace(self.h, problem_index, _ptr(out, f32p), cap, C.byref(n)), "ygz_sia_last_trace")
        return out[:n.value]

    def time(self, n_problems, iters):
        ms = C.c_float()
        _check(self.L.ygz_sia_time(self.h, n_problems, self.max_level, self.min_level, self.n_iter, self.eps, iters,
                                   C.byref(ms)), "ygz_sia_time")
        return ms.value
