"""ctypes binding of the CPU oracle (oracle/libygz_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

u8p = C.POINTER(C.c_uint8)
f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)
u32p = C.POINTER(C.c_uint32)


def _ptr(a, t):
    if a is None:
        return C.cast(None, t)
    return a.ctypes.data_as(t)


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


def build(native=False):
    target = "libygz_oracle_native.so" if native else "libygz_oracle.so"
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, target], check=True)
    return os.path.join(ORACLE_DIR, target)


_LIB = {}


def lib(native=False):
    if native in _LIB:
        return _LIB[native]
    target = os.path.join(ORACLE_DIR, "libygz_oracle_native.so" if native else "libygz_oracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".h", ".inc"))]
    if (not os.path.exists(target)) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in srcs):
        build(native)
    L = C.CDLL(target)
    L.orc_extractor_create.restype = C.c_void_p
    L.orc_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    L.orc_extractor_destroy.argtypes = [C.c_void_p]
    L.orc_extract.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, u8p, C.c_int]
    L.orc_extract_frame.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, u8p, C.c_int]
    L.orc_extract_frame_dso.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, u8p, C.c_int, i32p]
    L.orc_shi_tomasi.restype = C.c_float
    L.orc_shi_tomasi.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_level_size.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
    L.orc_get_level.argtypes = [C.c_void_p, C.c_int, C.c_int, u8p]
    L.orc_get_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.orc_get_level_keys.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.orc_extractor_tables.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p, i32p, i32p]
    L.orc_distribute_octtree.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_void_p, C.c_int]
    L.orc_extract_batch.restype = C.c_long
    L.orc_extract_batch.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_resize_linear_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int]
    L.orc_border_reflect101.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int]
    L.orc_gauss7_sigma2_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
    L.orc_fast9_16.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.orc_fast_atan2_array.argtypes = [f32p, f32p, f32p, C.c_int]
    L.orc_cosf_sinf_array.argtypes = [f32p, f32p, f32p, C.c_int]
    L.orc_hamming256.argtypes = [u8p, u8p]
    L.orc_search_by_bow.argtypes = [C.c_float, C.c_int, C.POINTER(FrameView), C.POINTER(MapPoints),
                                    C.POINTER(FrameView), i32p]
    L.orc_search_by_projection_points.argtypes = [C.c_float, C.POINTER(FrameView), C.POINTER(MapPoints), i32p,
                                                  C.c_int, C.c_float, C.c_int]
    L.orc_search_by_projection_last.argtypes = [C.c_int, C.POINTER(FrameView), C.POINTER(FrameView),
                                                C.POINTER(MapPoints), C.c_float, C.c_int, C.c_int]
    L.orc_search_by_projection_kf.argtypes = [C.c_int, C.POINTER(FrameView), C.POINTER(FrameView),
                                              C.POINTER(MapPoints), u8p, C.c_float, C.c_int]
    L.orc_search_for_initialization.argtypes = [C.c_float, C.c_int, C.POINTER(FrameView), C.POINTER(FrameView),
                                                f32p, i32p, C.c_int]
    L.orc_three_maxima.argtypes = [i32p, C.c_int, i32p, i32p, i32p]
    L.orc_search_by_projection_sim3.argtypes = [C.c_void_p, f32p, C.c_void_p, i32p, C.c_int, i32p, C.c_int]
    L.orc_fuse_sim3.argtypes = [C.c_void_p, f32p, C.c_void_p, i32p, C.c_int, C.c_float, i32p]
    L.orc_search_by_sim3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, i32p, i32p, C.c_float, f32p, f32p, C.c_float]
    L.orc_fuse_candidates.argtypes = [C.c_void_p, f32p, f32p, C.c_void_p, i32p, C.c_int, C.c_float, i32p, i32p]
    L.orc_fuse_candidates.restype = None
    L.orc_features_in_area.argtypes = [C.POINTER(FrameView), C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
                                       i32p, C.c_int]
    L.orc_bf_top2.argtypes = [u8p, C.c_int, u8p, C.c_int, i32p, i32p, i32p]
    L.orc_sia_run.argtypes = [C.POINTER(SiaProblem), C.c_int, C.c_int, C.c_int, C.c_float, f32p, f32p, f32p,
                              f32p, C.c_int, i32p]
    L.orc_find_direct_projection.argtypes = [C.c_void_p, C.c_int, C.c_void_p, f32p, f32p, f32p, C.c_float, C.c_int,
                                             C.c_void_p, C.c_int, f32p, i32p, u8p]
    L.orc_frontend_batch.restype = C.c_long
    L.orc_frontend_batch.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_float, C.c_int, C.POINTER(C.c_long), C.c_int]
    L.orc_se3_exp.argtypes = [f32p, f32p, f32p]
    L.orc_se3_mul.argtypes = [f32p, f32p, f32p, f32p, f32p, f32p]
    L.orc_se3_inv.argtypes = [f32p, f32p, f32p, f32p]
    _LIB[native] = L
    return L


class Extractor:
    """Oracle ORBextractor (restates src/ORBextractor.cc)."""

    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7, native=False):
        self.L = lib(native)
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = self.L.orc_extractor_create(nfeatures, scale, nlevels, ini_th, min_th)

    def __del__(self):
        try:
            self.L.orc_extractor_destroy(self.h)
        except Exception:
            pass

    def tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        fpl = np.zeros(n, np.int32)
        um = np.zeros(16, np.int32)
        self.L.orc_extractor_tables(self.h, _ptr(sc, f32p), _ptr(isc, f32p), _ptr(s2, f32p), _ptr(is2, f32p),
                                    _ptr(fpl, i32p), _ptr(um, i32p))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, features_per_level=fpl, umax=um)

    def extract(self, img, cap=None):
        img = np.ascontiguousarray(img, np.uint8)
        cap = cap or (self.nfeatures * 2 + 64)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.L.orc_extract(self.h, _ptr(img, u8p), img.shape[1], img.shape[0], img.strides[0],
                               kps.ctypes.data, _ptr(desc, u8p), cap)
        assert n >= 0, n
        return kps[:n].copy(), desc[:n].copy()

    def extract_frame(self, img, existing, cap=None):
        img = np.ascontiguousarray(img, np.uint8)
        ne = len(existing)
        cap = cap or (self.nfeatures * 2 + ne + 64)
        kps = np.zeros(cap, KP_DTYPE)
        kps[:ne] = existing
        desc = np.zeros((cap, 32), np.uint8)
        n = self.L.orc_extract_frame(self.h, _ptr(img, u8p), img.shape[1], img.shape[0], img.strides[0], ne,
                                     kps.ctypes.data, _ptr(desc, u8p), cap)
        assert n >= 0, n
        return kps[:n].copy(), desc[:n].copy()

    def extract_frame_dso(self, img, existing, grid_size=-1, cap=None):
        img = np.ascontiguousarray(img, np.uint8)
        ne = len(existing)
        cap = cap or (3 * (img.shape[0] // 7) * (img.shape[1] // 7) + ne)
        kps = np.zeros(cap, KP_DTYPE)
        kps[:ne] = existing
        desc = np.zeros((cap, 32), np.uint8)
        g = C.c_int(int(grid_size))
        n = self.L.orc_extract_frame_dso(self.h, _ptr(img, u8p), img.shape[1], img.shape[0], img.strides[0], ne,
                                         kps.ctypes.data, _ptr(desc, u8p), cap, C.byref(g))
        assert n >= 0, n
        return kps[:n].copy(), desc[:n].copy(), g.value

    def extract_tail(self, levels, blurred, cands, cap=None):
        """octree + IC_Angle + rBRIEF + scaling on externally produced levels / blurred levels / FAST candidates"""
        n = self.nlevels
        assert len(levels) == len(blurred) == len(cands) == n
        lv = [np.ascontiguousarray(x, np.uint8) for x in levels]
        bl = [np.ascontiguousarray(x, np.uint8) for x in blurred]
        cs = [np.ascontiguousarray(x, KP_DTYPE) for x in cands]
        PP = u8p * n
        w = (C.c_int * n)(*[x.shape[1] for x in lv]); h = (C.c_int * n)(*[x.shape[0] for x in lv])
        nc = (C.c_int * n)(*[len(x) for x in cs])
        cp = (C.c_void_p * n)(*[x.ctypes.data for x in cs])
        cap = cap or (self.nfeatures * 2 + 64)
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        self.L.orc_extract_tail.argtypes = [C.c_void_p, C.c_int, PP, PP, C.c_int * n, C.c_int * n, C.c_void_p * n, C.c_int * n,
                                            C.c_void_p, u8p, C.c_int]
        k = self.L.orc_extract_tail(self.h, n, PP(*[_ptr(x, u8p) for x in lv]), PP(*[_ptr(x, u8p) for x in bl]), w, h, cp, nc,
                                    kps.ctypes.data, _ptr(desc, u8p), cap)
        assert k >= 0, k
        return kps[:k].copy(), desc[:k].copy()

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        self.L.orc_level_size(self.h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def level(self, level, blurred=False):
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        self.L.orc_get_level(self.h, level, int(blurred), _ptr(out, u8p))
        return out

    def candidates(self, level, cap=200000):
        out = np.zeros(cap, KP_DTYPE)
        n = self.L.orc_get_candidates(self.h, level, out.ctypes.data, cap)
        assert n >= 0
        return out[:n].copy()

    def level_keys(self, level, cap=20000):
        out = np.zeros(cap, KP_DTYPE)
        n = self.L.orc_get_level_keys(self.h, level, out.ctypes.data, cap)
        assert n >= 0
        return out[:n].copy()

    def distribute(self, keys, minX, maxX, minY, maxY, N):
        keys = np.ascontiguousarray(keys, KP_DTYPE)
        out = np.zeros(max(len(keys), 1), KP_DTYPE)
        n = self.L.orc_distribute_octtree(self.h, keys.ctypes.data, len(keys), minX, maxX, minY, maxY, N,
                                          out.ctypes.data, len(out))
        assert n >= 0
        return out[:n].copy()


def resize(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_ptr(src, u8p), src.shape[1], src.shape[0], src.strides[0], _ptr(dst, u8p), dw, dh, dw)
    return dst


def border101(src, b):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((src.shape[0] + 2 * b, src.shape[1] + 2 * b), np.uint8)
    lib().orc_border_reflect101(_ptr(src, u8p), src.shape[1], src.shape[0], src.strides[0], _ptr(dst, u8p), b, dst.shape[1])
    return dst


def gauss7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros_like(src)
    lib().orc_gauss7_sigma2_u8(_ptr(src, u8p), src.shape[1], src.shape[0], src.strides[0], _ptr(dst, u8p), src.shape[1])
    return dst


def fast9(img, threshold, nms=True, cap=400000):
    assert img.dtype == np.uint8 and img.strides[1] == 1
    out = np.zeros(cap, KP_DTYPE)
    n = lib().orc_fast9_16(C.cast(img.ctypes.data, u8p), img.shape[1], img.shape[0], img.strides[0], threshold,
                           int(nms), out.ctypes.data, cap)
    assert n >= 0
    return out[:n].copy()


def fast_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros_like(y)
    lib().orc_fast_atan2_array(_ptr(y, f32p), _ptr(x, f32p), _ptr(out, f32p), len(y))
    return out


def cos_sin_deg(deg):
    deg = np.ascontiguousarray(deg, np.float32)
    c = np.zeros_like(deg)
    s = np.zeros_like(deg)
    lib().orc_cosf_sinf_array(_ptr(deg, f32p), _ptr(c, f32p), _ptr(s, f32p), len(deg))
    return c, s


def bf_top2(q, t):
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    bi = np.zeros(len(q), np.int32)
    bd = np.zeros(len(q), np.int32)
    sd = np.zeros(len(q), np.int32)
    lib().orc_bf_top2(_ptr(q, u8p), len(q), _ptr(t, u8p), len(t), _ptr(bi, i32p), _ptr(bd, i32p), _ptr(sd, i32p))
    return bi, bd, sd


def sia_wide_sums(on):
    """Test switch of the oracle aligner (calling thread): H / Jres / chi2 summed in double and rounded once."""
    lib().orc_sia_wide_sums(int(on))


def sia_run(problem, max_level, min_level, n_iter=10, eps=1e-6, trace_cap=192):
    """problem: _Keep(SiaProblem) built with cls=oracle_lib.SiaProblem.  Returns (n_tracked, q, t, H, trace)."""
    q = np.zeros(4, np.float32)
    t = np.zeros(3, np.float32)
    H = np.zeros(36, np.float32)
    tr = np.zeros((trace_cap, 8), np.float32)
    tn = C.c_int()
    n = lib().orc_sia_run(problem.ptr(), max_level, min_level, n_iter, eps, _ptr(q, f32p), _ptr(t, f32p), _ptr(H, f32p),
                          _ptr(tr, f32p), trace_cap, C.byref(tn))
    return n, q, t, H.reshape(6, 6), tr[:tn.value]


def se3_exp(x6):
    x6 = np.ascontiguousarray(x6, np.float32)
    q = np.zeros(4, np.float32)
    t = np.zeros(3, np.float32)
    lib().orc_se3_exp(_ptr(x6, f32p), _ptr(q, f32p), _ptr(t, f32p))
    return q, t


def frontend_batch(imgs, threads, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7, nnratio=0.7, check_ori=True,
                   native=True, repeat=1, vocab=None, levelsup=4):
    """CPU baseline: extract (+ Frame::ComputeBoW through `vocab`, a Vocabulary of the SAME library build) + SearchByBoW(prev
    frame) over imgs[n,h,w]; returns (matches, keypoints)."""
    imgs = np.ascontiguousarray(imgs, np.uint8)
    n, h, w = imgs.shape
    kp = C.c_long()
    L = lib(native)
    L.orc_frontend_batch_voc.restype = C.c_long
    L.orc_frontend_batch_voc.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_float, C.c_int, C.POINTER(C.c_long), C.c_int, C.c_void_p, C.c_int]
    m = L.orc_frontend_batch_voc(nfeatures, scale, nlevels, ini_th, min_th, _ptr(imgs, u8p), n, w, h, threads,
                                 nnratio, int(check_ori), C.byref(kp), repeat, vocab.h if vocab is not None else None, levelsup)
    return m, kp.value


def bow_match_pair(a, b, vocab, levelsup=4, nnratio=0.7, check_ori=True):
    """Frame::ComputeBoW of both frames + SearchByBoW(KF = a, F = b); a, b = (keypoints, descriptors); returns nmatches"""
    (ka, da), (kb, db) = a, b
    L = vocab.L_
    L.orc_bow_match_pair.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, u8p, C.c_int, C.c_void_p, u8p, C.c_int]
    return L.orc_bow_match_pair(vocab.h, levelsup, nnratio, int(check_ori), ka.ctypes.data, _ptr(da, u8p), len(ka),
                                kb.ctypes.data, _ptr(db, u8p), len(kb))


def find_direct_projection(ref_frames, cur_frame, K, scale, inv_scale, inv_level_sigma2_1, items):
    """Oracle FindDirectProjection over a batch.  ref_frames: list of _Keep(DirectFrame) (the product's ctypes struct has the
    same layout as orc_direct_frame), cur_frame: _Keep(DirectFrame), items: structured array with the layout of
    ygz_direct_item / orc_direct_item.  Returns (px[n,2], level[n], ok[n])."""
    L = lib()
    n = len(items)
    items = np.ascontiguousarray(items)
    Frame = type(cur_frame.struct)
    refs = (Frame * len(ref_frames))(*[f.struct for f in ref_frames])
    K = np.ascontiguousarray(K, np.float32)
    scale = np.ascontiguousarray(scale, np.float32)
    inv_scale = np.ascontiguousarray(inv_scale, np.float32)
    px = np.zeros((n, 2), np.float32); lvl = np.zeros(n, np.int32); ok = np.zeros(n, np.uint8)
    rc = L.orc_find_direct_projection(C.cast(refs, C.c_void_p), len(ref_frames), C.cast(C.pointer(cur_frame.struct), C.c_void_p),
                                      K.ctypes.data_as(f32p), scale.ctypes.data_as(f32p), inv_scale.ctypes.data_as(f32p),
                                      float(inv_level_sigma2_1), len(scale), items.ctypes.data, n,
                                      px.ctypes.data_as(f32p), lvl.ctypes.data_as(i32p), ok.ctypes.data_as(u8p))
    assert rc == 0, rc
    return px, lvl, ok


def compute_stereo_matches(pyrL, pyrR, keysL, descL, keysR, descR, scale, inv_scale, mb, mbf):
    """Oracle Frame::ComputeStereoMatches.  pyrL / pyrR: _Keep(DirectFrame) built by make_direct_frame (poses unused)."""
    L = lib()
    L.orc_compute_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                             C.c_void_p, C.c_int, f32p, f32p, C.c_float, C.c_float, f32p, f32p]
    keysL = np.ascontiguousarray(keysL, KP_DTYPE); keysR = np.ascontiguousarray(keysR, KP_DTYPE)
    descL = np.ascontiguousarray(descL, np.uint8); descR = np.ascontiguousarray(descR, np.uint8)
    scale = np.ascontiguousarray(scale, np.float32); inv_scale = np.ascontiguousarray(inv_scale, np.float32)
    u = np.zeros(len(keysL), np.float32); z = np.zeros(len(keysL), np.float32)
    n = L.orc_compute_stereo_matches(C.cast(C.pointer(pyrL.struct), C.c_void_p), C.cast(C.pointer(pyrR.struct), C.c_void_p), len(scale),
                                     keysL.ctypes.data, descL.ctypes.data, len(keysL), keysR.ctypes.data, descR.ctypes.data, len(keysR),
                                     scale.ctypes.data_as(f32p), inv_scale.ctypes.data_as(f32p), float(mb), float(mbf),
                                     u.ctypes.data_as(f32p), z.ctypes.data_as(f32p))
    return u, z, n


class Vocabulary:
    """Oracle DBoW2 vocabulary (oracle/vocabulary.cpp)."""

    def __init__(self, path=None, data=None, native=False):
        self.L_ = lib(native)
        L = self.L_
        L.orc_vocab_load.restype = C.c_void_p; L.orc_vocab_load.argtypes = [C.c_char_p]
        L.orc_vocab_from_memory.restype = C.c_void_p; L.orc_vocab_from_memory.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_vocab_destroy.argtypes = [C.c_void_p]
        L.orc_vocab_info.argtypes = [C.c_void_p] + [i32p] * 6
        L.orc_vocab_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, i32p, f32p, i32p]
        L.orc_vocab_transform_bow.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, i32p, C.POINTER(C.c_double), i32p, i32p, i32p, i32p]
        if path is not None:
            self.h = L.orc_vocab_load(str(path).encode())
        else:
            buf = np.ascontiguousarray(np.frombuffer(data, np.uint8))
            self.h = L.orc_vocab_from_memory(buf.ctypes.data, buf.size)
        assert self.h, "vocabulary did not load"
        v = [C.c_int() for _ in range(6)]
        L.orc_vocab_info(self.h, *[C.byref(x) for x in v])
        self.k, self.L, self.scoring, self.weighting, self.n_nodes, self.n_words = [x.value for x in v]

    def __del__(self):
        try:
            self.L_.orc_vocab_destroy(self.h)
        except Exception:
            pass

    def transform_features(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        wid = np.zeros(n, np.int32); w = np.zeros(n, np.float32); nid = np.zeros(n, np.int32)
        self.L_.orc_vocab_transform(self.h, desc.ctypes.data, n, levelsup, wid.ctypes.data_as(i32p), w.ctypes.data_as(f32p), nid.ctypes.data_as(i32p))
        return wid, w, nid

    def transform(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        ids = np.zeros(max(n, 1), np.int32); vals = np.zeros(max(n, 1), np.float64)
        fvn = C.c_int(); node = np.zeros(max(n, 1), np.int32); off = np.zeros(n + 2, np.int32); idx = np.zeros(max(n, 1), np.int32)
        nb = self.L_.orc_vocab_transform_bow(self.h, desc.ctypes.data, n, levelsup, ids.ctypes.data_as(i32p),
                                            vals.ctypes.data_as(C.POINTER(C.c_double)), C.byref(fvn), node.ctypes.data_as(i32p),
                                            off.ctypes.data_as(i32p), idx.ctypes.data_as(i32p))
        fv = {int(node[j]): [int(x) for x in idx[off[j]:off[j + 1]]] for j in range(fvn.value)}
        return ids[:nb].copy(), vals[:nb].copy(), fv
