"""ctypes binding of oracle/_ref/libextractor_ref.so: the REFERENCE's own ygz::ORBextractor (src/ORBextractor.cc compiled
unmodified against oracle/ref_mock).  TEST INFRASTRUCTURE ONLY — the pin of the oracle restatement."""
import ctypes as C
import os
import subprocess
import numpy as np

from oracle_lib import KP_DTYPE, ORACLE_DIR, _ptr, f32p, u8p

class RefThrows(Exception):
    """The reference raised a C++ exception (e.g. std::length_error from vpIniNodes.resize(nIni < 0), src/ORBextractor.cc:545)."""


REF_SRC = "/root/reference/src/ORBextractor.cc"
_LIBS = {}


def _load(name):
    if name in _LIBS:
        return _LIBS[name]
    path = os.path.join(ORACLE_DIR, "_ref", name)
    if os.path.exists(REF_SRC):                 # build container: (re)build from the reference where it lies
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "_ref/" + name], check=True)
    if not os.path.exists(path):
        _LIBS[name] = None
        return None
    _LIBS[name] = C.CDLL(path)
    return _LIBS[name]


def extractor_lib():
    L = _load("libextractor_ref.so")
    if L is None:
        return None
    L.ref_orb_create.restype = C.c_void_p
    L.ref_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    L.ref_orb_destroy.argtypes = [C.c_void_p]
    L.ref_orb_tables.argtypes = [C.c_void_p, C.c_int, f32p]
    L.ref_orb_extract.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, u8p, C.c_int]
    L.ref_orb_frame_pyramid.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_int]
    L.ref_orb_level.argtypes = [C.c_void_p, C.c_int, C.c_int, u8p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.ref_orb_extract_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, u8p, C.c_int]
    L.ref_arena.argtypes = [C.c_int]
    return L


class RefExtractor:
    """The reference's ygz::ORBextractor (include/ORBextractor.h:45-172)."""
    ORBSLAM_KEYPOINT, FAST_KEYPOINT, DSO_KEYPOINT = 0, 1, 2

    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7, creation_order=True):
        self.L = extractor_lib()
        assert self.L is not None, "oracle/_ref/libextractor_ref.so is missing"
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.creation_order = creation_order
        self.h = self.L.ref_orb_create(nfeatures, scale, nlevels, ini_th, min_th)

    def __del__(self):
        try:
            self.L.ref_orb_destroy(self.h)
        except Exception:
            pass

    def tables(self):
        out = {}
        for i, k in enumerate(("scale", "inv_scale", "sigma2", "inv_sigma2")):
            v = np.zeros(self.nlevels, np.float32)
            self.L.ref_orb_tables(self.h, i, _ptr(v, f32p))
            out[k] = v
        return out

    def extract(self, img, cap=None):
        """operator()(image, mask, keypoints, descriptors)"""
        img = np.ascontiguousarray(img, np.uint8)
        cap = cap or (self.nfeatures * 2 + 64)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        self.L.ref_arena(int(self.creation_order))
        n = self.L.ref_orb_extract(self.h, _ptr(img, u8p), img.shape[1], img.shape[0], img.strides[0],
                                   kps.ctypes.data, _ptr(desc, u8p), cap)
        if n == -2:
            raise RefThrows()
        assert 0 <= n <= cap, n
        return kps[:n].copy(), desc[:n].copy()

    def level(self, level, border=0):
        w, h = C.c_int(), C.c_int()
        assert self.L.ref_orb_level(self.h, level, border, None, C.byref(w), C.byref(h)) == 0
        out = np.zeros((h.value, w.value), np.uint8)
        self.L.ref_orb_level(self.h, level, border, _ptr(out, u8p), C.byref(w), C.byref(h))
        return out

    def extract_frame(self, img, existing, method=0, cap=None):
        """Frame ctor pyramid (src/Frame.cc:807-813) + operator()(Frame*, frame->mvKeys, desc, method, leftEye=true)"""
        img = np.ascontiguousarray(img, np.uint8)
        existing = np.ascontiguousarray(existing, KP_DTYPE)
        ne = len(existing)
        cap = cap or (3 * (img.shape[0] // 7) * (img.shape[1] // 7) + self.nfeatures * 2 + ne + 64)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        self.L.ref_arena(int(self.creation_order))
        self.L.ref_orb_frame_pyramid(self.h, _ptr(img, u8p), img.shape[1], img.shape[0], img.strides[0])
        n = self.L.ref_orb_extract_frame(self.h, method, existing.ctypes.data, ne, kps.ctypes.data, _ptr(desc, u8p), cap)
        if n == -2:
            raise RefThrows()
        assert 0 <= n <= cap, n
        return kps[:n].copy(), desc[:n].copy()


def matcher_lib():
    """oracle/_ref/libmatcher_ref.so: the reference's own ygz::ORBmatcher (src/ORBmatcher.cc + src/Align.cc compiled unmodified
    against oracle/ref_mock; oracle/ref_matcher_shim.cpp).  Entry points mirror the oracle's orc_* matcher signatures."""
    L = _load("libmatcher_ref.so")
    if L is None:
        return None
    i32p = C.POINTER(C.c_int)
    L.refm_hamming256.argtypes = [u8p, u8p]
    L.refm_three_maxima.argtypes = [i32p, C.c_int, i32p, i32p, i32p]
    L.refm_search_by_bow.argtypes = [C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, i32p]
    L.refm_search_by_projection_points.argtypes = [C.c_float, C.c_void_p, C.c_void_p, i32p, C.c_int, C.c_float, C.c_int]
    L.refm_search_by_projection_last.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int]
    L.refm_search_by_projection_kf.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, u8p, C.c_float, C.c_int]
    L.refm_search_for_initialization.argtypes = [C.c_float, C.c_int, C.c_void_p, C.c_void_p, f32p, i32p, C.c_int]
    L.refm_search_by_bow_kf.argtypes = [C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, i32p]
    L.refm_search_for_triangulation.argtypes = [C.c_int, C.c_void_p, C.c_void_p, f32p, C.POINTER(C.c_float), C.POINTER(C.c_float), f32p,
                                                C.c_int, i32p]
    return L


def find_direct_projection(ref_frames, cur_frame, K, scale, inv_scale, inv_level_sigma2_1, items):
    """The reference's ORBmatcher::FindDirectProjection per item (refm_find_direct_projection); arguments as
    oracle_lib.find_direct_projection."""
    L = matcher_lib()
    i32p = C.POINTER(C.c_int)
    n = len(items)
    items = np.ascontiguousarray(items)
    Frame = type(cur_frame.struct)
    refs = (Frame * len(ref_frames))(*[f.struct for f in ref_frames])
    K = np.ascontiguousarray(K, np.float32)
    scale = np.ascontiguousarray(scale, np.float32)
    inv_scale = np.ascontiguousarray(inv_scale, np.float32)
    px = np.zeros((n, 2), np.float32); lvl = np.zeros(n, np.int32); ok = np.zeros(n, np.uint8)
    L.refm_find_direct_projection.argtypes = [C.c_void_p, C.c_int, C.c_void_p, f32p, f32p, f32p, C.c_float, C.c_int, C.c_void_p, C.c_int,
                                              f32p, i32p, u8p]
    rc = L.refm_find_direct_projection(C.cast(refs, C.c_void_p), len(ref_frames), C.cast(C.pointer(cur_frame.struct), C.c_void_p),
                                       K.ctypes.data_as(f32p), scale.ctypes.data_as(f32p), inv_scale.ctypes.data_as(f32p),
                                       float(inv_level_sigma2_1), len(scale), items.ctypes.data, n,
                                       px.ctypes.data_as(f32p), lvl.ctypes.data_as(i32p), ok.ctypes.data_as(u8p))
    assert rc == 0, rc
    return px, lvl, ok


class RefVocabulary:
    """The reference's ygz::ORBVocabulary (DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>), oracle/ref_vocab_shim.cpp.
    Same methods as oracle_lib.Vocabulary.  loadFromBinaryFile only reads files: `data` goes through a temporary file."""

    def __init__(self, path=None, data=None):
        import tempfile
        self.L_ = _load("libvocab_ref.so")
        assert self.L_ is not None, "oracle/_ref/libvocab_ref.so is missing"
        L = self.L_
        i32p = C.POINTER(C.c_int)
        L.refv_load.restype = C.c_void_p; L.refv_load.argtypes = [C.c_char_p]
        L.refv_destroy.argtypes = [C.c_void_p]
        L.refv_info.argtypes = [C.c_void_p] + [i32p] * 6
        L.refv_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, i32p, f32p, i32p]
        L.refv_transform_bow.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, i32p, C.POINTER(C.c_double), i32p, i32p, i32p, i32p]
        if path is None:
            with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
                f.write(bytes(data))
                tmp = f.name
            try:
                self.h = L.refv_load(tmp.encode())
            finally:
                os.unlink(tmp)
        else:
            self.h = L.refv_load(str(path).encode())
        assert self.h, "vocabulary did not load"
        v = [C.c_int() for _ in range(6)]
        L.refv_info(self.h, *[C.byref(x) for x in v])
        self.k, self.L, self.scoring, self.weighting, self.n_nodes, self.n_words = [x.value for x in v]

    def __del__(self):
        try:
            self.L_.refv_destroy(self.h)
        except Exception:
            pass

    def transform_features(self, desc, levelsup=4):
        i32p = C.POINTER(C.c_int)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        wid = np.zeros(n, np.int32); w = np.zeros(n, np.float32); nid = np.zeros(n, np.int32)
        self.L_.refv_transform(self.h, desc.ctypes.data, n, levelsup, wid.ctypes.data_as(i32p), w.ctypes.data_as(f32p), nid.ctypes.data_as(i32p))
        return wid, w, nid

    def transform(self, desc, levelsup=4):
        i32p = C.POINTER(C.c_int)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        ids = np.zeros(max(n, 1), np.int32); vals = np.zeros(max(n, 1), np.float64)
        fvn = C.c_int(); node = np.zeros(max(n, 1), np.int32); off = np.zeros(n + 2, np.int32); idx = np.zeros(max(n, 1), np.int32)
        nb = self.L_.refv_transform_bow(self.h, desc.ctypes.data, n, levelsup, ids.ctypes.data_as(i32p), vals.ctypes.data_as(C.POINTER(C.c_double)),
                                        C.byref(fvn), node.ctypes.data_as(i32p), off.ctypes.data_as(i32p), idx.ctypes.data_as(i32p))
        fv = {int(node[j]): [int(x) for x in idx[off[j]:off[j + 1]]] for j in range(fvn.value)}
        return ids[:nb].copy(), vals[:nb].copy(), fv


def adapters_lib():
    """oracle/_ref/libadapters_e2e.so: oracle/ref_matcher_shim.cpp compiled with -DSHIM_USE_ADAPTERS — the same caller code as
    libmatcher_ref.so, but ygz::ORBmatcher / ORBextractor / SparseImgAlign are the product's literal-signature adapters
    (include/ygz/adapters.h) over libygz_b200.so.  Needs a GPU at run time.  Built in the build container only (it needs the
    reference's Common.h to compile); travels prebuilt."""
    path = os.path.join(ORACLE_DIR, "_ref", "libadapters_e2e.so")
    if os.path.exists(REF_SRC):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "_ref/libadapters_e2e.so"], check=True)
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    i32p = C.POINTER(C.c_int)
    L.adp_hamming256.argtypes = [u8p, u8p]
    L.adp_search_by_bow.argtypes = [C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, i32p]
    L.adp_search_by_projection_points.argtypes = [C.c_float, C.c_void_p, C.c_void_p, i32p, C.c_int, C.c_float, C.c_int]
    L.adp_search_by_projection_last.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int]
    L.adp_search_by_projection_kf.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, u8p, C.c_float, C.c_int]
    L.adp_search_for_initialization.argtypes = [C.c_float, C.c_int, C.c_void_p, C.c_void_p, f32p, i32p, C.c_int]
    L.adp_search_by_bow_kf.argtypes = [C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, i32p]
    L.adp_search_for_triangulation.argtypes = [C.c_int, C.c_void_p, C.c_void_p, f32p, C.POINTER(C.c_float), C.POINTER(C.c_float), f32p,
                                               C.c_int, i32p]
    L.adp_extract_image.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, u8p, C.c_int]
    L.adp_extract_frame.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, u8p, C.c_int, C.c_int, u8p]
    L.adp_sia_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, f32p, f32p]
    L.adp_search_by_projection_sim3.argtypes = [C.c_void_p, f32p, C.c_void_p, i32p, C.c_int, i32p, C.c_int]
    L.adp_fuse_sim3.argtypes = [C.c_void_p, f32p, C.c_void_p, i32p, C.c_int, C.c_float, i32p]
    L.adp_search_by_sim3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, i32p, i32p, C.c_float, f32p, f32p, C.c_float]
    L.adp_fuse_kf.argtypes = [C.c_void_p, f32p, C.c_void_p, i32p, C.c_int, i32p, C.c_int, C.c_float, u8p, i32p, i32p, i32p]
    return L


def sia_run(problem, max_level, min_level, n_iter=10):
    """The reference's ygz::SparseImgAlign::run (oracle/_ref/libsia_ref.so: src/SparseImageAlign.cc + NLSSolver_impl.hpp compiled
    unmodified; oracle/ref_sia_shim.cpp).  problem: _Keep(SiaProblem) built with cls=oracle_lib.SiaProblem.
    Returns (n_tracked, q, t, H[6, 6], chi2) or None when the library is not built."""
    L = _load("libsia_ref.so")
    if L is None:
        return None
    L.refs_sia_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, C.POINTER(C.c_float)]
    q = np.zeros(4, np.float32); t = np.zeros(3, np.float32); H = np.zeros(36, np.float32)
    chi2 = C.c_float()
    n = L.refs_sia_run(C.cast(problem.ptr(), C.c_void_p), max_level, min_level, n_iter, _ptr(q, f32p), _ptr(t, f32p), _ptr(H, f32p), C.byref(chi2))
    return n, q, t, H.reshape(6, 6), chi2.value


def compute_stereo_matches(pyrL, pyrR, keysL, descL, keysR, descR, scale, inv_scale, mb, mbf):
    """The reference's own Frame::ComputeStereoMatches (oracle/_ref/libstereo_ref.so, oracle/ref_stereo_shim.cpp); arguments and
    returns as oracle_lib.compute_stereo_matches.  None when the library is not built."""
    L = _load("libstereo_ref.so")
    if L is None:
        return None
    L.refst_compute_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                               f32p, f32p, C.c_float, C.c_float, f32p, f32p]
    keysL = np.ascontiguousarray(keysL, KP_DTYPE); keysR = np.ascontiguousarray(keysR, KP_DTYPE)
    descL = np.ascontiguousarray(descL, np.uint8); descR = np.ascontiguousarray(descR, np.uint8)
    scale = np.ascontiguousarray(scale, np.float32); inv_scale = np.ascontiguousarray(inv_scale, np.float32)
    u = np.zeros(len(keysL), np.float32); z = np.zeros(len(keysL), np.float32)
    n = L.refst_compute_stereo_matches(C.cast(C.pointer(pyrL.struct), C.c_void_p), C.cast(C.pointer(pyrR.struct), C.c_void_p), len(scale),
                                       keysL.ctypes.data, descL.ctypes.data, len(keysL), keysR.ctypes.data, descR.ctypes.data, len(keysR),
                                       _ptr(scale, f32p), _ptr(inv_scale, f32p), float(mb), float(mbf), _ptr(u, f32p), _ptr(z, f32p))
    return u, z, n


def matcher_sim3_lib():
    """The Sim3-family entry points of oracle/_ref/libmatcher_ref.so (SearchByProjection(KF, Scw), Fuse(KF, Scw), SearchBySim3)."""
    L = matcher_lib()
    if L is None:
        return None
    i32p = C.POINTER(C.c_int)
    L.refm_search_by_projection_sim3.argtypes = [C.c_void_p, f32p, C.c_void_p, i32p, C.c_int, i32p, C.c_int]
    L.refm_fuse_sim3.argtypes = [C.c_void_p, f32p, C.c_void_p, i32p, C.c_int, C.c_float, i32p]
    L.refm_search_by_sim3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, i32p, i32p, C.c_float, f32p, f32p, C.c_float]
    u8p = C.POINTER(C.c_uint8)
    L.refm_fuse_kf.argtypes = [C.c_void_p, f32p, C.c_void_p, i32p, C.c_int, i32p, C.c_int, C.c_float, u8p, i32p, i32p, i32p]
    L.refm_fuse_kf_each.argtypes = [C.c_void_p, f32p, C.c_void_p, i32p, C.c_int, C.c_float, i32p]
    L.refm_fuse_kf_each.restype = None
    return L
