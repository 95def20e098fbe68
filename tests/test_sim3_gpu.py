"""GPU parity of the Sim3 family (SearchByProjection(KF, Scw, ...), Fuse(KF, Scw, ...), SearchBySim3: src/ORBmatcher.cc:265-373, 888-1004,
1006-1216) through the C ABI against the oracle (pinned to the reference's compiled code by tests/test_ref_sim3.py): index arrays and counts."""
import ctypes as C
import numpy as np
import pytest

import helpers as H
import test_ref_sim3 as S

pytestmark = pytest.mark.gpu
K, BOUNDS = S.K, S.BOUNDS


@pytest.fixture(scope="module")
def frames(oracle, synth):
    return H.two_frames(oracle, synth, seed=1)


@pytest.mark.parametrize("s,th", [(1.0, 10), (1.3, 10)])
def test_projection_and_fuse_sim3(gpu, oracle, frames, s, th):
    (_, ka, da), (_, kb, db) = frames
    rng = np.random.default_rng(11)
    R = S.rot((0.2, 1.0, 0.1), 0.03); t = np.array([0.05, -0.02, 0.1])
    Xw, nrm, mind, maxd = S.world_points(ka, R, t, rng, shift=(-9.0, -5.0))
    m, extra = len(ka), 40
    Xw2 = np.concatenate([Xw, np.zeros((extra, 3), np.float32)]); nrm2 = np.concatenate([nrm, np.zeros((extra, 3), np.float32)])
    desc = np.concatenate([da, rng.integers(0, 256, (extra, 32), dtype=np.uint8)])
    bad = (rng.random(m + extra) < 0.05).astype(np.uint8)
    kwargs = dict(bad=bad, min_dist=np.r_[mind, np.ones(extra)], max_dist=np.r_[maxd, np.ones(extra)], normal=nrm2)
    Scw = np.eye(4, dtype=np.float32); Scw[:3, :3] = (s * R).astype(np.float32); Scw[:3, 3] = (s * t).astype(np.float32)
    ids = rng.permutation(m).astype(np.int32)
    slots = np.full(len(kb), -1, np.int32)
    slots[::7] = m + rng.integers(0, extra, len(slots[::7]))
    mt = gpu.ORBmatcher(0.75, True)
    res = {}
    for name, fcls, mcls in (("gpu", gpu.FrameView, gpu.MapPoints), ("orc", oracle.FrameView, oracle.MapPoints)):
        mps = gpu.make_map_points(Xw2, desc, cls=mcls, **kwargs)
        kf = gpu.make_frame_view(kb, db, K=K, bounds=BOUNDS, scale_factors=H.SCALE8, cls=fcls)
        matched = slots.copy()
        kf2 = gpu.make_frame_view(kb, db, map_point=slots.copy(), K=K, bounds=BOUNDS, scale_factors=H.SCALE8, cls=fcls)
        rep = np.full(m, -1, np.int32)
        if name == "gpu":
            n1 = gpu.lib().ygz_search_by_projection_sim3(mt.h, kf.ptr(), Scw.ctypes.data_as(gpu.f32p), mps.ptr(), ids.ctypes.data_as(gpu.i32p), m,
                                                         matched.ctypes.data_as(gpu.i32p), th)
            n2 = gpu.lib().ygz_fuse_sim3(mt.h, kf2.ptr(), Scw.ctypes.data_as(gpu.f32p), mps.ptr(), ids.ctypes.data_as(gpu.i32p), m, float(th) / 2,
                                         rep.ctypes.data_as(gpu.i32p))
        else:
            L = oracle.lib()
            n1 = L.orc_search_by_projection_sim3(S.vp(kf), Scw.ctypes.data_as(oracle.f32p), S.vp(mps), ids.ctypes.data_as(oracle.i32p), m,
                                                 matched.ctypes.data_as(oracle.i32p), th)
            n2 = L.orc_fuse_sim3(S.vp(kf2), Scw.ctypes.data_as(oracle.f32p), S.vp(mps), ids.ctypes.data_as(oracle.i32p), m, C.c_float(float(th) / 2),
                                 rep.ctypes.data_as(oracle.i32p))
        res[name] = (n1, matched, n2, rep, kf2.map_point.copy())
    g, o = res["gpu"], res["orc"]
    assert g[0] == o[0] and (g[1] == o[1]).all() and g[0] > 50
    assert g[2] == o[2] and (g[3] == o[3]).all() and (g[4] == o[4]).all() and g[2] > 50


@pytest.mark.parametrize("s12,th", [(1.0, 7.5), (1.15, 7.5)])
def test_search_by_sim3(gpu, oracle, frames, s12, th):
    (_, ka, da), (_, kb, db) = frames
    rng = np.random.default_rng(12)
    R2 = S.rot((0.1, 1.0, 0.0), 0.01); t2 = np.array([-0.1, -0.05, 0.02])
    n1, n2 = len(ka), len(kb)
    X1, nr1, mi1, ma1 = S.world_points(ka, np.eye(3), np.zeros(3), rng)
    X2, nr2, mi2, ma2 = S.world_points(kb, R2, t2, rng)
    Xw = np.concatenate([X1, X2]); nrm = np.concatenate([nr1, nr2])
    kwargs = dict(bad=(rng.random(n1 + n2) < 0.04).astype(np.uint8), min_dist=np.r_[mi1, mi2] * 0.5, max_dist=np.r_[ma1, ma2] * 1.5, normal=nrm)
    mp1 = np.where(rng.random(n1) < 0.85, np.arange(n1), -1).astype(np.int32)
    mp2 = np.where(rng.random(n2) < 0.85, n1 + np.arange(n2), -1).astype(np.int32)
    R12 = R2.T.astype(np.float32); t12 = (-(R2.T @ t2) * s12).astype(np.float32)
    pre = np.full(n1, -1, np.int32)
    pre[::13] = n1 + rng.integers(0, n2, len(pre[::13]))
    idx2 = np.full(n1 + n2, -1, np.int32); idx2[n1:] = np.arange(n2)
    mt = gpu.ORBmatcher(0.75, True)
    res = {}
    for name, fcls, mcls in (("gpu", gpu.FrameView, gpu.MapPoints), ("orc", oracle.FrameView, oracle.MapPoints)):
        mps = gpu.make_map_points(Xw, np.concatenate([da, db]), cls=mcls, **kwargs)
        k1 = gpu.make_frame_view(ka, da, map_point=mp1, K=K, bounds=BOUNDS, scale_factors=H.SCALE8, cls=fcls)
        k2 = gpu.make_frame_view(kb, db, map_point=mp2, K=K, bounds=BOUNDS, scale_factors=H.SCALE8, Rcw=R2.astype(np.float32), tcw=t2.astype(np.float32), cls=fcls)
        m12 = pre.copy()
        if name == "gpu":
            n = gpu.lib().ygz_search_by_sim3(mt.h, k1.ptr(), k2.ptr(), mps.ptr(), m12.ctypes.data_as(gpu.i32p), idx2.ctypes.data_as(gpu.i32p), float(s12),
                                             R12.reshape(9).ctypes.data_as(gpu.f32p), t12.ctypes.data_as(gpu.f32p), float(th))
        else:
            n = oracle.lib().orc_search_by_sim3(S.vp(k1), S.vp(k2), S.vp(mps), m12.ctypes.data_as(oracle.i32p), idx2.ctypes.data_as(oracle.i32p),
                                                C.c_float(s12), R12.reshape(9).ctypes.data_as(oracle.f32p), t12.ctypes.data_as(oracle.f32p), C.c_float(th))
        res[name] = (n, m12)
    assert res["gpu"][0] == res["orc"][0] and (res["gpu"][1] == res["orc"][1]).all()
    if s12 == 1.0:
        assert res["gpu"][0] > 30
