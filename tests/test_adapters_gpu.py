"""End-to-end run of the literal-signature adapters (include/ygz/adapters.h) on the GPU.

oracle/_ref/libadapters_e2e.so is oracle/ref_matcher_shim.cpp — the caller code that drives the REFERENCE's compiled
ORBmatcher.cc in tests/test_ref_matcher.py — compiled a second time with ygz::ORBmatcher / ORBextractor / SparseImgAlign bound to
the product's adapters: Frame / KeyFrame / MapPoint objects in, `matcher.SearchByBoW(pKF, F, vpMapPointMatches)` etc. called with
the reference's signatures, the pointer-graph gather done by the adapters, libygz_b200.so underneath.  Results must equal the
oracle's (which tests/test_ref_*.py pin to the reference's own compiled code)."""
import ctypes as C
import numpy as np
import pytest

import helpers as H
import ref_lib
import test_ref_matcher as R
from test_ref_matcher import vp, op_, K

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def adp(gpu):
    L = ref_lib.adapters_lib()
    if L is None:
        pytest.skip("oracle/_ref/libadapters_e2e.so not built")
    return L


@pytest.fixture(scope="module")
def frames(oracle, synth):
    return H.two_frames(oracle, synth, seed=1)


@pytest.fixture(scope="module")
def frames_small(oracle, synth):
    return {s: H.two_frames(oracle, synth, seed=s, w=480, h=360, nfeatures=700) for s in (2, 5)}


def test_search_by_bow_both(adp, pkg, oracle, frames, frames_small):
    (_, ka, da), (_, kb, db) = frames
    rng = np.random.default_rng(2)
    mp = np.arange(len(ka), dtype=np.int32)
    mp[rng.random(len(ka)) < 0.2] = -1
    bad = (rng.random(len(ka)) < 0.05).astype(np.uint8)
    kf = pkg.make_frame_view(ka, da, map_point=mp, node_of_feature=H.synth_vocab_nodes(da), cls=oracle.FrameView)
    f = pkg.make_frame_view(kb, db, node_of_feature=H.synth_vocab_nodes(db), cls=oracle.FrameView)
    mps = pkg.make_map_points(np.zeros((len(ka), 3)), da, bad=bad, cls=oracle.MapPoints)
    o, a = np.zeros(len(kb), np.int32), np.zeros(len(kb), np.int32)
    on = oracle.lib().orc_search_by_bow(0.75, 1, kf.ptr(), mps.ptr(), f.ptr(), o.ctypes.data_as(oracle.i32p))
    an = adp.adp_search_by_bow(0.75, 1, vp(kf), vp(mps), vp(f), a.ctypes.data_as(oracle.i32p))
    assert on == an and (o == a).all() and on > 50
    # KeyFrame-to-KeyFrame
    (_, ka, da), (_, kb, db) = frames_small[2]
    M = len(ka) + len(kb)
    mp1 = np.where(rng.random(len(ka)) < 0.8, np.arange(len(ka)), -1).astype(np.int32)
    mp2 = np.where(rng.random(len(kb)) < 0.8, len(ka) + np.arange(len(kb)), -1).astype(np.int32)
    bad = (rng.random(M) < 0.05).astype(np.uint8)
    f1 = pkg.make_frame_view(ka, da, map_point=mp1, node_of_feature=H.synth_vocab_nodes(da), cls=oracle.FrameView)
    f2 = pkg.make_frame_view(kb, db, map_point=mp2, node_of_feature=H.synth_vocab_nodes(db), cls=oracle.FrameView)
    mps = pkg.make_map_points(np.zeros((M, 3)), np.concatenate([da, db]), bad=bad, cls=oracle.MapPoints)
    o12, a12 = np.zeros(len(ka), np.int32), np.zeros(len(ka), np.int32)
    Lo = oracle.lib()
    Lo.orc_search_by_bow_kf.argtypes = [C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, oracle.i32p]
    on = Lo.orc_search_by_bow_kf(0.8, 1, vp(f1), vp(f2), vp(mps), o12.ctypes.data_as(oracle.i32p))
    an = adp.adp_search_by_bow_kf(0.8, 1, vp(f1), vp(f2), vp(mps), a12.ctypes.data_as(oracle.i32p))
    assert on == an and (o12 == a12).all() and on > 50


@pytest.mark.parametrize("th,check_level,stereo", [(1.0, True, False), (5.0, True, True)])
def test_search_by_projection_points(adp, pkg, oracle, frames, th, check_level, stereo):
    (_, ka, da), (_, kb, db) = frames
    rng = np.random.default_rng(3)
    wp, track, obs, z = R._projection_setup(ka, rng)
    ur = (kb["x"] - 30.0).astype(np.float32) if stereo else None
    if stereo:
        ur[::3] = -1
    pre = np.full(len(kb), -1, np.int32)
    pre[::11] = rng.integers(0, len(ka), len(pre[::11]))
    bad = (rng.random(len(ka)) < 0.03).astype(np.uint8)
    q = rng.permutation(len(ka)).astype(np.int32)
    res = []
    for fn, cv in ((oracle.lib().orc_search_by_projection_points, op_), (adp.adp_search_by_projection_points, vp)):
        f = pkg.make_frame_view(kb, db, map_point=pre.copy(), u_right=ur, K=K, mbf=40.0, mb=0.09, scale_factors=H.SCALE8, cls=oracle.FrameView)
        mps = pkg.make_map_points(wp, da, bad=bad, observations=obs, track=track, cls=oracle.MapPoints)
        n = fn(C.c_float(0.8), cv(f), cv(mps), q.ctypes.data_as(oracle.i32p), len(q), C.c_float(th), int(check_level))
        res.append((n, f.map_point.copy()))
    assert res[0][0] == res[1][0] and (res[0][1] == res[1][1]).all() and res[0][0] > 20


@pytest.mark.parametrize("mono,check_level,tz", [(True, True, 0.0), (False, True, 0.3)])
def test_search_by_projection_last_and_kf(adp, pkg, oracle, frames, mono, check_level, tz):
    (_, ka, da), (_, kb, db) = frames
    rng = np.random.default_rng(4)
    wp, track, obs, z = R._projection_setup(ka, rng)
    mp_last = np.arange(len(ka), dtype=np.int32)
    mp_last[rng.random(len(ka)) < 0.3] = -1
    outl = (rng.random(len(ka)) < 0.05).astype(np.uint8)
    tl = np.array([0.0, 0.0, tz], np.float32)
    ur = np.where(rng.random(len(kb)) < 0.5, kb["x"] - 8.0, -1).astype(np.float32) if not mono else None
    found = (rng.random(len(ka)) < 0.1).astype(np.uint8)
    bad = (rng.random(len(ka)) < 0.03).astype(np.uint8)
    maxd, mind = (z * 1.5).astype(np.float32), (z * 0.5).astype(np.float32)
    out = []
    for L, pre, cv in ((oracle.lib(), "orc_", op_), (adp, "adp_", vp)):
        cur = pkg.make_frame_view(kb, db, u_right=ur, K=K, mb=0.09, mbf=40.0, scale_factors=H.SCALE8, cls=oracle.FrameView)
        last = pkg.make_frame_view(ka, da, map_point=mp_last, outlier=outl, K=K, tcw=tl, scale_factors=H.SCALE8, cls=oracle.FrameView)
        mps = pkg.make_map_points(wp, da, bad=bad, observations=obs, min_dist=mind, max_dist=maxd, cls=oracle.MapPoints)
        cur2 = pkg.make_frame_view(kb, db, K=K, scale_factors=H.SCALE8, cls=oracle.FrameView)
        kfv = pkg.make_frame_view(ka, da, map_point=mp_last, K=K, scale_factors=H.SCALE8, cls=oracle.FrameView)
        n1 = getattr(L, pre + "search_by_projection_last")(1, cv(cur), cv(last), cv(mps), C.c_float(15.0), int(mono), int(check_level))
        n2 = getattr(L, pre + "search_by_projection_kf")(1, cv(cur2), cv(kfv), cv(mps), found.ctypes.data_as(oracle.u8p), C.c_float(10.0), 100)
        out.append((n1, cur.map_point.copy(), n2, cur2.map_point.copy()))
    o, a = out
    assert o[0] == a[0] and (o[1] == a[1]).all() and o[2] == a[2] and (o[3] == a[3]).all()
    assert o[0] > 20 and o[2] > 20


def test_initialization_and_triangulation(adp, pkg, oracle, frames, frames_small):
    (_, ka, da), (_, kb, db) = frames
    prev = np.stack([ka["x"], ka["y"]], -1).astype(np.float32)
    res = []
    for fn, cv in ((oracle.lib().orc_search_for_initialization, op_), (adp.adp_search_for_initialization, vp)):
        f1 = pkg.make_frame_view(ka, da, K=K, scale_factors=H.SCALE8, cls=oracle.FrameView)
        f2 = pkg.make_frame_view(kb, db, K=K, scale_factors=H.SCALE8, cls=oracle.FrameView)
        p = prev.copy()
        m12 = np.zeros(len(ka), np.int32)
        n = fn(C.c_float(0.9), 1, cv(f1), cv(f2), p.ctypes.data_as(oracle.f32p), m12.ctypes.data_as(oracle.i32p), 100)
        res.append((n, m12, p))
    assert res[0][0] == res[1][0] and (res[0][1] == res[1][1]).all() and (res[0][2] == res[1][2]).all() and res[0][0] > 20
    (_, ka, da), (_, kb, db) = frames_small[5]
    rng = np.random.default_rng(5)
    mp1 = np.where(rng.random(len(ka)) < 0.3, np.arange(len(ka)), -1).astype(np.int32)
    mp2 = np.where(rng.random(len(kb)) < 0.3, np.arange(len(kb)), -1).astype(np.int32)
    ur1 = np.where(rng.random(len(ka)) < 0.5, ka["x"] - 10, -1).astype(np.float32)
    ur2 = np.where(rng.random(len(kb)) < 0.5, kb["x"] - 10, -1).astype(np.float32)
    dx, dy = -9.0, -5.0
    F12 = np.array([[0, 0, -dy], [0, 0, dx], [dy, -dx, 0]], np.float32)
    t2 = np.array([40 * dx / K[0], 40 * dy / K[1], 1.0], np.float32)
    f1 = pkg.make_frame_view(ka, da, map_point=mp1, u_right=ur1, K=K, node_of_feature=H.synth_vocab_nodes(da), bounds=(0, 0, 480, 360), cls=oracle.FrameView)
    f2 = pkg.make_frame_view(kb, db, map_point=mp2, u_right=ur2, K=K, tcw=t2, node_of_feature=H.synth_vocab_nodes(db), bounds=(0, 0, 480, 360),
                             cls=oracle.FrameView)
    ls2 = ((1.2 ** np.arange(8)).astype(np.float32) ** 2).astype(np.float32)
    ap, op = np.zeros((len(ka), 2), np.int32), np.zeros((len(ka), 2), np.int32)
    ex, ey = C.c_float(), C.c_float()
    an = adp.adp_search_for_triangulation(1, vp(f1), vp(f2), F12.reshape(9).ctypes.data_as(oracle.f32p), C.byref(ex), C.byref(ey),
                                          ls2.ctypes.data_as(oracle.f32p), 0, ap.ctypes.data_as(oracle.i32p))
    Lo = oracle.lib()
    Lo.orc_search_for_triangulation.argtypes = [C.c_int, C.c_void_p, C.c_void_p, oracle.f32p, C.c_float, C.c_float, oracle.f32p, C.c_int, oracle.i32p]
    on = Lo.orc_search_for_triangulation(1, vp(f1), vp(f2), F12.reshape(9).ctypes.data_as(oracle.f32p), ex.value, ey.value,
                                         ls2.ctypes.data_as(oracle.f32p), 0, op.ctypes.data_as(oracle.i32p))
    assert on == an and (op[:on] == ap[:an]).all() and on > 20


@pytest.mark.parametrize("w,h,seed", [(752, 480, 1), (640, 360, 3)])
def test_extractor_adapters(adp, oracle, synth, w, h, seed):
    img = synth.synth_frame(seed, w, h)
    ox = oracle.Extractor(1000, 1.2, 8, 20, 7)
    ok, od = ox.extract(img)
    cap = 4096
    kps = np.zeros(cap, oracle.KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    n = adp.adp_extract_image(1000, 1.2, 8, 20, 7, img.ctypes.data_as(oracle.u8p), w, h, img.strides[0], kps.ctypes.data,
                              desc.ctypes.data_as(oracle.u8p), cap)
    assert n == len(ok) and kps[:n].tobytes() == ok.tobytes() and (desc[:n] == od).all()
    # Frame* overload: pre-existing keys first (their descriptors recomputed), new keys appended; public mvImagePyramid filled
    existing = ok[::40].copy()
    ek, ed = ox.extract_frame(img, existing)
    kps[:len(existing)] = existing
    lw, lh = ox.level_size(3)
    probe = np.zeros((lh, lw), np.uint8)
    n = adp.adp_extract_frame(1000, 1.2, 8, 20, 7, img.ctypes.data_as(oracle.u8p), w, h, img.strides[0], 0, len(existing), kps.ctypes.data,
                              desc.ctypes.data_as(oracle.u8p), cap, 3, probe.ctypes.data_as(oracle.u8p))
    assert n == len(ek) and kps[:n].tobytes() == ek.tobytes() and (desc[:n] == ed).all()
    assert (probe == ox.level(3)).all()


def test_sia_adapter(adp, gpu, oracle, synth):
    import test_sia_gpu as S
    po, s = S.build_problem(gpu, oracle, synth, 2, oracle.SiaProblem)
    n, q, t, H_, trace = oracle.sia_run(po, 3, 1, 10, 1e-6)
    aq, at = np.zeros(4, np.float32), np.zeros(3, np.float32)
    an = adp.adp_sia_run(C.cast(po.ptr(), C.c_void_p), 3, 1, 10, aq.ctypes.data_as(oracle.f32p), at.ctypes.data_as(oracle.f32p))
    assert an == n and n > 150
    assert S.pose_diff(q, t, aq, at) < 1e-4
    assert S.pose_diff(aq, at, s["q_true"], s["t_true"]) < 5e-3


def test_sim3_family_adapters(adp, pkg, oracle, frames):
    """SearchByProjection(KeyFrame*, cv::Mat Scw, ...), Fuse(KeyFrame*, cv::Mat Scw, ...), SearchBySim3 with the reference's signatures
    (KeyFrame* / MapPoint* / cv::Mat in, AddObservation + AddMapPoint performed on the objects) vs the oracle."""
    import test_ref_sim3 as S
    (_, ka, da), (_, kb, db) = frames
    rng = np.random.default_rng(11)
    Rm = S.rot((0.2, 1.0, 0.1), 0.03); t = np.array([0.05, -0.02, 0.1])
    Xw, nrm, mind, maxd = S.world_points(ka, Rm, t, rng, shift=(-9.0, -5.0))
    m, extra, s, th = len(ka), 40, 1.3, 10
    Xw2 = np.concatenate([Xw, np.zeros((extra, 3), np.float32)]); nrm2 = np.concatenate([nrm, np.zeros((extra, 3), np.float32)])
    desc = np.concatenate([da, rng.integers(0, 256, (extra, 32), dtype=np.uint8)])
    mps = pkg.make_map_points(Xw2, desc, bad=(rng.random(m + extra) < 0.05).astype(np.uint8), min_dist=np.r_[mind, np.ones(extra)],
                              max_dist=np.r_[maxd, np.ones(extra)], normal=nrm2, cls=oracle.MapPoints)
    Scw = np.eye(4, dtype=np.float32); Scw[:3, :3] = (s * Rm).astype(np.float32); Scw[:3, 3] = (s * t).astype(np.float32)
    ids = rng.permutation(m).astype(np.int32)
    slots = np.full(len(kb), -1, np.int32)
    slots[::7] = m + rng.integers(0, extra, len(slots[::7]))
    out = []
    for L, pre in ((oracle.lib(), "orc_"), (adp, "adp_")):
        kf = pkg.make_frame_view(kb, db, K=S.K, bounds=S.BOUNDS, scale_factors=H.SCALE8, cls=oracle.FrameView)
        matched = slots.copy()
        n1 = getattr(L, pre + "search_by_projection_sim3")(vp(kf), Scw.ctypes.data_as(oracle.f32p), vp(mps), ids.ctypes.data_as(oracle.i32p), m,
                                                            matched.ctypes.data_as(oracle.i32p), th)
        kf2 = pkg.make_frame_view(kb, db, map_point=slots.copy(), K=S.K, bounds=S.BOUNDS, scale_factors=H.SCALE8, cls=oracle.FrameView)
        rep = np.full(m, -1, np.int32)
        n2 = getattr(L, pre + "fuse_sim3")(vp(kf2), Scw.ctypes.data_as(oracle.f32p), vp(mps), ids.ctypes.data_as(oracle.i32p), m, C.c_float(5.0),
                                           rep.ctypes.data_as(oracle.i32p))
        out.append((n1, matched, n2, rep, kf2.map_point.copy()))
    o, a = out
    assert o[0] == a[0] and (o[1] == a[1]).all() and o[2] == a[2] and (o[3] == a[3]).all() and (o[4] == a[4]).all() and o[0] > 50 and o[2] > 50
    # SearchBySim3
    rng = np.random.default_rng(12)
    R2 = S.rot((0.1, 1.0, 0.0), 0.01); t2 = np.array([-0.1, -0.05, 0.02])
    n1, n2 = len(ka), len(kb)
    X1, nr1, mi1, ma1 = S.world_points(ka, np.eye(3), np.zeros(3), rng)
    X2, nr2, mi2, ma2 = S.world_points(kb, R2, t2, rng)
    mps = pkg.make_map_points(np.concatenate([X1, X2]), np.concatenate([da, db]), bad=(rng.random(n1 + n2) < 0.04).astype(np.uint8),
                              min_dist=np.r_[mi1, mi2] * 0.5, max_dist=np.r_[ma1, ma2] * 1.5, normal=np.concatenate([nr1, nr2]), cls=oracle.MapPoints)
    mp1 = np.where(rng.random(n1) < 0.85, np.arange(n1), -1).astype(np.int32)
    mp2 = np.where(rng.random(n2) < 0.85, n1 + np.arange(n2), -1).astype(np.int32)
    R12 = R2.T.astype(np.float32); t12 = (-(R2.T @ t2)).astype(np.float32)
    pre12 = np.full(n1, -1, np.int32)
    pre12[::13] = n1 + rng.integers(0, n2, len(pre12[::13]))
    idx2 = np.full(n1 + n2, -1, np.int32); idx2[n1:] = np.arange(n2)
    out = []
    for L, pre in ((oracle.lib(), "orc_"), (adp, "adp_")):
        k1 = pkg.make_frame_view(ka, da, map_point=mp1, K=S.K, bounds=S.BOUNDS, scale_factors=H.SCALE8, cls=oracle.FrameView)
        k2 = pkg.make_frame_view(kb, db, map_point=mp2, K=S.K, bounds=S.BOUNDS, scale_factors=H.SCALE8, Rcw=R2.astype(np.float32), tcw=t2.astype(np.float32),
                                 cls=oracle.FrameView)
        m12 = pre12.copy()
        n = getattr(L, pre + "search_by_sim3")(vp(k1), vp(k2), vp(mps), m12.ctypes.data_as(oracle.i32p), idx2.ctypes.data_as(oracle.i32p), C.c_float(1.0),
                                               R12.reshape(9).ctypes.data_as(oracle.f32p), t12.ctypes.data_as(oracle.f32p), C.c_float(7.5))
        out.append((n, m12))
    assert out[0][0] == out[1][0] and (out[0][1] == out[1][1]).all() and out[0][0] > 30
