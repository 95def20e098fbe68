"""ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:748-886) on the GPU:
  * ygz_fuse_candidates == orc_fuse_candidates (best index and distance per candidate point, bit for bit; tests/test_ref_fuse.py pins the
    oracle to the reference's compiled code);
  * the literal-signature adapter ORBmatcher::Fuse(pKF, vpMapPoints, th) (include/ygz/adapters.h, search on the device + upstream's graph
    mutation replayed on the caller's objects) leaves the map in the state the REFERENCE's own compiled Fuse leaves it in
    (oracle/_ref/libmatcher_ref.so, prebuilt): slots of both key frames, bad flags, observation counts, mpReplaced, return value."""
import ctypes as C
import numpy as np
import pytest

import fuse_scene as FS
import helpers as H
import ref_lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frames(oracle, synth):
    return H.two_frames(oracle, synth, seed=1)


@pytest.mark.parametrize("stereo,th,seed", [(False, 3.0, 21), (True, 3.0, 22), (True, 6.0, 23), (False, 1.5, 24)])
def test_fuse_candidates_match_oracle(gpu, pkg, oracle, frames, stereo, th, seed):
    (_, ka, da), (_, kb, db) = frames
    S = FS.build(ka, da, kb, db, seed, stereo)
    ids = S["ids"]
    n = len(ids)
    out = []
    for cls_mp, cls_fv, call in ((oracle.MapPoints, oracle.FrameView, "oracle"), (None, None, "gpu")):
        mps = pkg.make_map_points(S["Xw"], S["desc"], bad=S["bad"], min_dist=S["mind"], max_dist=S["maxd"], normal=S["normal"], cls=cls_mp)
        kf = pkg.make_frame_view(kb, db, u_right=S["u_right"], K=FS.K, mbf=FS.MBF, bounds=FS.BOUNDS, scale_factors=S["sf"], Rcw=S["R"], tcw=S["t"],
                                 cls=cls_fv)
        bi = np.zeros(n, np.int32); bd = np.zeros(n, np.int32)
        f32p, i32p = oracle.f32p, oracle.i32p
        if call == "oracle":
            oracle.lib().orc_fuse_candidates(C.cast(kf.ptr(), C.c_void_p), S["Ow"].ctypes.data_as(f32p), S["inv_sigma2"].ctypes.data_as(f32p),
                                             C.cast(mps.ptr(), C.c_void_p), ids.ctypes.data_as(i32p), n, C.c_float(th), bi.ctypes.data_as(i32p),
                                             bd.ctypes.data_as(i32p))
        else:
            m = gpu.ORBmatcher(0.6, True)
            rc = gpu.lib().ygz_fuse_candidates(m.h, kf.ptr(), S["Ow"].ctypes.data_as(f32p), S["inv_sigma2"].ctypes.data_as(f32p), mps.ptr(),
                                               ids.ctypes.data_as(i32p), n, C.c_float(th), bi.ctypes.data_as(i32p), bd.ctypes.data_as(i32p))
            assert rc == 0, gpu.lib().ygz_last_error()
        out.append((bi, bd))
    assert (out[0][0] == out[1][0]).all() and (out[0][1] == out[1][1]).all()
    assert (out[0][0] >= 0).sum() > 300 and (out[0][0] < 0).sum() > 40


def test_fuse_candidates_edge_cases(gpu, pkg, oracle, frames):
    (_, ka, da), (_, kb, db) = frames
    S = FS.build(ka, da, kb, db, 25, False)
    mps = pkg.make_map_points(S["Xw"], S["desc"], bad=S["bad"], min_dist=S["mind"], max_dist=S["maxd"], normal=S["normal"])
    kf = pkg.make_frame_view(kb, db, K=FS.K, mbf=FS.MBF, bounds=FS.BOUNDS, scale_factors=S["sf"], Rcw=S["R"], tcw=S["t"])
    m = gpu.ORBmatcher(0.6, True)
    f32p, i32p = oracle.f32p, oracle.i32p
    L = gpu.lib()
    # empty list, all-NULL list
    bi = np.zeros(4, np.int32); bd = np.zeros(4, np.int32)
    assert L.ygz_fuse_candidates(m.h, kf.ptr(), S["Ow"].ctypes.data_as(f32p), S["inv_sigma2"].ctypes.data_as(f32p), mps.ptr(), None, 0, C.c_float(3.0),
                                 bi.ctypes.data_as(i32p), bd.ctypes.data_as(i32p)) == 0
    ids = np.full(4, -1, np.int32)
    assert L.ygz_fuse_candidates(m.h, kf.ptr(), S["Ow"].ctypes.data_as(f32p), S["inv_sigma2"].ctypes.data_as(f32p), mps.ptr(), ids.ctypes.data_as(i32p), 4,
                                 C.c_float(3.0), bi.ctypes.data_as(i32p), bd.ctypes.data_as(i32p)) == 0
    assert (bi == -1).all() and (bd == 256).all()
    # index outside the table, table without normals: argument errors, not crashes
    ids[0] = S["m"] + 5
    assert L.ygz_fuse_candidates(m.h, kf.ptr(), S["Ow"].ctypes.data_as(f32p), S["inv_sigma2"].ctypes.data_as(f32p), mps.ptr(), ids.ctypes.data_as(i32p), 4,
                                 C.c_float(3.0), bi.ctypes.data_as(i32p), bd.ctypes.data_as(i32p)) < 0
    bare = pkg.make_map_points(S["Xw"], S["desc"], min_dist=S["mind"], max_dist=S["maxd"])
    ids[0] = 0
    assert L.ygz_fuse_candidates(m.h, kf.ptr(), S["Ow"].ctypes.data_as(f32p), S["inv_sigma2"].ctypes.data_as(f32p), bare.ptr(), ids.ctypes.data_as(i32p), 4,
                                 C.c_float(3.0), bi.ctypes.data_as(i32p), bd.ctypes.data_as(i32p)) < 0


@pytest.mark.parametrize("stereo,th,seed", [(False, 3.0, 21), (True, 3.0, 22), (True, 5.0, 26)])
def test_fuse_adapter_leaves_the_reference_map_state(gpu, pkg, oracle, frames, stereo, th, seed):
    adp = ref_lib.adapters_lib()
    ref = ref_lib.matcher_sim3_lib()
    if adp is None or ref is None:
        pytest.skip("oracle/_ref/libadapters_e2e.so / libmatcher_ref.so not built")
    (_, ka, da), (_, kb, db) = frames
    S = FS.build(ka, da, kb, db, seed, stereo)
    a = FS.run_full(adp.adp_fuse_kf, pkg, oracle, kb, db, S, th)
    r = FS.run_full(ref.refm_fuse_kf, pkg, oracle, kb, db, S, th)
    assert a[0] == r[0] and a[0] > 150
    for x, y, what in zip(a[1:], r[1:], ("key-frame slots", "bad", "observations", "replaced_by", "second key frame")):
        assert (x == y).all(), what
