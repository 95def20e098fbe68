"""GPU parity of the key-frame-to-key-frame searches against the oracle: ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...)
(src/ORBmatcher.cc:480-595) and ORBmatcher::SearchForTriangulation (:597-746 with CheckDistEpipolarLine :136-153).
Index arrays identical (host candidate enumeration -> one device distance launch -> host greedy in reference order)."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _views(gpu, oracle, ka, da, kb, db, mp1, mp2, ur1=None, ur2=None):
    out = []
    for cls in (None, oracle.FrameView):
        f1 = gpu.make_frame_view(ka, da, map_point=mp1, u_right=ur1, node_of_feature=H.synth_vocab_nodes(da), cls=cls)
        f2 = gpu.make_frame_view(kb, db, map_point=mp2, u_right=ur2, node_of_feature=H.synth_vocab_nodes(db), cls=cls)
        out.append((f1, f2))
    return out


@pytest.mark.parametrize("seed,ratio,ori", [(1, 0.8, True), (2, 0.6, True), (3, 0.9, False)])
def test_search_by_bow_kf(gpu, oracle, synth, seed, ratio, ori):
    (_, ka, da), (_, kb, db) = H.two_frames(oracle, synth, seed=seed, w=480, h=360, nfeatures=700)
    rng = np.random.default_rng(seed)
    M = len(ka) + len(kb)
    mp1 = np.where(rng.random(len(ka)) < 0.8, np.arange(len(ka)), -1).astype(np.int32)
    mp2 = np.where(rng.random(len(kb)) < 0.8, len(ka) + np.arange(len(kb)), -1).astype(np.int32)
    bad = (rng.random(M) < 0.05).astype(np.uint8)
    (g1, g2), (o1, o2) = _views(gpu, oracle, ka, da, kb, db, mp1, mp2)
    gm = gpu.make_map_points(np.zeros((M, 3)), np.concatenate([da, db]), bad=bad)
    om = gpu.make_map_points(np.zeros((M, 3)), np.concatenate([da, db]), bad=bad, cls=oracle.MapPoints)
    mt = gpu.ORBmatcher(ratio, ori)
    n, m12 = mt.SearchByBoWKF(g1, g2, gm)
    o12 = np.zeros(len(ka), np.int32)
    import ctypes as C
    L = oracle.lib()
    L.orc_search_by_bow_kf.argtypes = [C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, oracle.i32p]
    on = L.orc_search_by_bow_kf(ratio, int(ori), C.cast(o1.ptr(), C.c_void_p), C.cast(o2.ptr(), C.c_void_p), C.cast(om.ptr(), C.c_void_p),
                                o12.ctypes.data_as(oracle.i32p))
    assert n == on and (m12 == o12).all()
    assert n > 50 and (m12[m12 >= 0] >= len(ka)).all()


@pytest.mark.parametrize("seed,only_stereo", [(4, False), (5, True), (6, False)])
def test_search_for_triangulation(gpu, oracle, synth, seed, only_stereo):
    import ctypes as C
    shift = (9, 5)
    (_, ka, da), (_, kb, db) = H.two_frames(oracle, synth, seed=seed, shift=shift, w=480, h=360, nfeatures=700)
    rng = np.random.default_rng(seed)
    mp1 = np.where(rng.random(len(ka)) < 0.3, np.arange(len(ka)), -1).astype(np.int32)       # most features are untracked
    mp2 = np.where(rng.random(len(kb)) < 0.3, np.arange(len(kb)), -1).astype(np.int32)
    ur1 = np.where(rng.random(len(ka)) < 0.5, ka["x"] - 10, -1).astype(np.float32)
    ur2 = np.where(rng.random(len(kb)) < 0.5, kb["x"] - 10, -1).astype(np.float32)
    (g1, g2), (o1, o2) = _views(gpu, oracle, ka, da, kb, db, mp1, mp2, ur1, ur2)
    dx, dy = -float(shift[0]), -float(shift[1])                   # a -> b image motion: the epipolar lines are parallel to it
    F12 = np.array([[0, 0, -dy], [0, 0, dx], [dy, -dx, 0]], np.float32)
    epi = (240.0 + 40 * dx, 180.0 + 40 * dy)                      # a far epipole in that direction
    ls2 = (1.2 ** np.arange(8)).astype(np.float32) ** 2
    mt = gpu.ORBmatcher(0.6, True)
    n, pairs = mt.SearchForTriangulation(g1, g2, F12, epi, ls2, only_stereo)
    opairs = np.zeros((len(ka), 2), np.int32)
    L = oracle.lib()
    L.orc_search_for_triangulation.argtypes = [C.c_int, C.c_void_p, C.c_void_p, oracle.f32p, C.c_float, C.c_float, oracle.f32p, C.c_int, oracle.i32p]
    on = L.orc_search_for_triangulation(1, C.cast(o1.ptr(), C.c_void_p), C.cast(o2.ptr(), C.c_void_p), F12.reshape(9).ctypes.data_as(oracle.f32p),
                                        epi[0], epi[1], ls2.ctypes.data_as(oracle.f32p), int(only_stereo), opairs.ctypes.data_as(oracle.i32p))
    assert n == on and (pairs == opairs[:on]).all()
    assert n > 20
    d = np.stack([kb["x"][pairs[:, 1]] - ka["x"][pairs[:, 0]], kb["y"][pairs[:, 1]] - ka["y"][pairs[:, 0]]], -1)
    assert (np.abs(d - np.array([dx, dy])).max(1) < 3).mean() > 0.8          # the accepted pairs follow the true motion
    assert (mp1[pairs[:, 0]] < 0).all() and (mp2[pairs[:, 1]] < 0).all()
    if only_stereo:
        assert (ur1[pairs[:, 0]] >= 0).all() and (ur2[pairs[:, 1]] >= 0).all()
