"""The oracle's restatement of the Sim3 family of ORBmatcher (loop closing: SearchByProjection(KF, Scw, ...) src/ORBmatcher.cc:265-373,
Fuse(KF, Scw, ...) :888-1004, SearchBySim3 :1006-1216) against the REFERENCE's own compiled code (oracle/_ref/libmatcher_ref.so): index
arrays and counts identical.  These functions do their geometry in CV_32F cv::Mat algebra; OpenCV is an absent third-party dependency, so
both sides share the stand-in arithmetic of oracle/ref_mock/mock_cv.hpp (products accumulate in double).  Build container only."""
import ctypes as C
import numpy as np
import pytest

import helpers as H
import ref_lib

K = (458.654, 457.296, 367.215, 248.375)
BOUNDS = (0.0, 0.0, 752.0, 480.0)


@pytest.fixture(scope="module")
def ref():
    L = ref_lib.matcher_sim3_lib()
    if L is None:
        pytest.skip("oracle/_ref/libmatcher_ref.so not built (no /root/reference here)")
    return L


@pytest.fixture(scope="module")
def frames(oracle, synth):
    return H.two_frames(oracle, synth, seed=1)


def vp(x):
    return C.cast(x.ptr(), C.c_void_p)


def rot(ax, ang):
    ax = np.asarray(ax, np.float64); ax /= np.linalg.norm(ax)
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return (np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx)


def world_points(keys, R, t, rng, shift=(0.0, 0.0), noise=0.8):
    """Points that project (through Xc = R Xw + t) onto keys + shift (+ noise), with scale-invariance distances centred on the key's octave
    and normals pointing back at the camera."""
    fx, fy, cx, cy = K
    n = len(keys)
    z = rng.uniform(2, 10, n)
    u = keys["x"].astype(np.float64) + shift[0] + rng.normal(0, noise, n); v = keys["y"].astype(np.float64) + shift[1] + rng.normal(0, noise, n)
    Xc = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], -1)
    Xw = (Xc - t) @ R                                       # R^T (Xc - t)
    Ow = -R.T @ t
    PO = Xw - Ow
    dist = np.linalg.norm(PO, axis=1)
    normal = PO / dist[:, None]
    flip = rng.random(n) < 0.05
    normal[flip] *= -1                                      # viewing-angle rejections
    maxd = dist * 1.2 ** (keys["octave"] + 0.5)
    mind = dist * 0.4
    far = rng.random(n) < 0.05
    mind[far] = dist[far] * 2.0                             # depth outside the invariance region
    return Xw.astype(np.float32), normal.astype(np.float32), mind.astype(np.float32), maxd.astype(np.float32)


@pytest.mark.parametrize("s,th", [(1.0, 10), (1.3, 10), (0.8, 4)])
def test_search_by_projection_and_fuse_sim3(ref, pkg, oracle, frames, s, th):
    (_, ka, da), (_, kb, db) = frames
    rng = np.random.default_rng(11)
    R = rot((0.2, 1.0, 0.1), 0.03); t = np.array([0.05, -0.02, 0.1])
    Xw, nrm, mind, maxd = world_points(ka, R, t, rng, shift=(-9.0, -5.0))
    m = len(ka)
    extra = 40                                               # map points already sitting in KF slots (some bad)
    Xw2 = np.concatenate([Xw, np.zeros((extra, 3), np.float32)]); nrm2 = np.concatenate([nrm, np.zeros((extra, 3), np.float32)])
    desc = np.concatenate([da, rng.integers(0, 256, (extra, 32), dtype=np.uint8)])
    bad = (rng.random(m + extra) < 0.05).astype(np.uint8)
    mps = pkg.make_map_points(Xw2, desc, bad=bad, min_dist=np.r_[mind, np.ones(extra)], max_dist=np.r_[maxd, np.ones(extra)], normal=nrm2,
                              cls=oracle.MapPoints)
    Scw = np.eye(4, dtype=np.float32); Scw[:3, :3] = (s * R).astype(np.float32); Scw[:3, 3] = (s * t).astype(np.float32)
    ids = rng.permutation(m).astype(np.int32)
    slots = np.full(len(kb), -1, np.int32)
    slots[::7] = m + rng.integers(0, extra, len(slots[::7]))
    # ---- SearchByProjection(KF, Scw, vpPoints, vpMatched, th)
    res = []
    for fn in (oracle.lib().orc_search_by_projection_sim3, ref.refm_search_by_projection_sim3):
        kf = pkg.make_frame_view(kb, db, K=K, bounds=BOUNDS, scale_factors=H.SCALE8, cls=oracle.FrameView)
        matched = slots.copy()
        fn.restype = C.c_int
        n = fn(vp(kf), Scw.ctypes.data_as(oracle.f32p), vp(mps), ids.ctypes.data_as(oracle.i32p), m, matched.ctypes.data_as(oracle.i32p), th)
        res.append((n, matched))
    assert res[0][0] == res[1][0] and (res[0][1] == res[1][1]).all()
    assert res[0][0] > 50
    # ---- Fuse(KF, Scw, vpPoints, th, vpReplacePoint)
    res = []
    for fn in (oracle.lib().orc_fuse_sim3, ref.refm_fuse_sim3):
        kf = pkg.make_frame_view(kb, db, map_point=slots.copy(), K=K, bounds=BOUNDS, scale_factors=H.SCALE8, cls=oracle.FrameView)
        rep = np.full(m, -1, np.int32)
        fn.restype = C.c_int
        n = fn(vp(kf), Scw.ctypes.data_as(oracle.f32p), vp(mps), ids.ctypes.data_as(oracle.i32p), m, C.c_float(float(th) / 2), rep.ctypes.data_as(oracle.i32p))
        res.append((n, rep, kf.map_point.copy()))
    assert res[0][0] == res[1][0] and (res[0][1] == res[1][1]).all() and (res[0][2] == res[1][2]).all()
    assert res[0][0] > 50 and (res[0][1] >= 0).any() and (res[0][2] != slots).any()        # both branches taken: replace candidates and new slots


@pytest.mark.parametrize("s12,th", [(1.0, 7.5), (1.15, 7.5), (1.0, 3.0)])
def test_search_by_sim3(ref, pkg, oracle, frames, s12, th):
    (_, ka, da), (_, kb, db) = frames
    rng = np.random.default_rng(12)
    # world = camera 1; camera 2: Xc2 = R2 Xw + t2 (a small motion that moves image content by roughly (-9, -5) px at the scene depth)
    R2 = rot((0.1, 1.0, 0.0), 0.01); t2 = np.array([-0.1, -0.05, 0.02])
    n1, n2 = len(ka), len(kb)
    X1, nr1, mi1, ma1 = world_points(ka, np.eye(3), np.zeros(3), rng)
    X2, nr2, mi2, ma2 = world_points(kb, R2, t2, rng)
    Xw = np.concatenate([X1, X2]); nrm = np.concatenate([nr1, nr2])
    mps = pkg.make_map_points(Xw, np.concatenate([da, db]), bad=(rng.random(n1 + n2) < 0.04).astype(np.uint8),
                              min_dist=np.r_[mi1, mi2] * 0.5, max_dist=np.r_[ma1, ma2] * 1.5, normal=nrm, cls=oracle.MapPoints)
    mp1 = np.where(rng.random(n1) < 0.85, np.arange(n1), -1).astype(np.int32)
    mp2 = np.where(rng.random(n2) < 0.85, n1 + np.arange(n2), -1).astype(np.int32)
    R12 = R2.T.astype(np.float32); t12 = (-(R2.T @ t2) * s12).astype(np.float32)
    pre = np.full(n1, -1, np.int32)
    pre[::13] = n1 + rng.integers(0, n2, len(pre[::13]))                  # matches found earlier (by BoW)
    idx2 = np.full(n1 + n2, -1, np.int32)
    idx2[n1:] = np.arange(n2)                                              # KF2's own points are observed in KF2 at their feature index
    res = []
    for fn in (oracle.lib().orc_search_by_sim3, ref.refm_search_by_sim3):
        k1 = pkg.make_frame_view(ka, da, map_point=mp1, K=K, bounds=BOUNDS, scale_factors=H.SCALE8, cls=oracle.FrameView)
        k2 = pkg.make_frame_view(kb, db, map_point=mp2, K=K, bounds=BOUNDS, scale_factors=H.SCALE8, Rcw=R2.astype(np.float32),
                                 tcw=t2.astype(np.float32), cls=oracle.FrameView)
        m12 = pre.copy()
        fn.restype = C.c_int
        n = fn(vp(k1), vp(k2), vp(mps), m12.ctypes.data_as(oracle.i32p), idx2.ctypes.data_as(oracle.i32p), C.c_float(s12),
               R12.reshape(9).ctypes.data_as(oracle.f32p), t12.ctypes.data_as(oracle.f32p), C.c_float(th))
        res.append((n, m12))
    assert res[0][0] == res[1][0] and (res[0][1] == res[1][1]).all()
    if s12 == 1.0 and th > 5:
        assert res[0][0] > 30
