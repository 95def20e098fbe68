"""GPU parity of the batched, device-resident projection searches (csrc/projection_batch.cuh) against the oracle run pair by pair:
ORBmatcher::SearchByProjection(Current, Last, th, bMono, checkLevel) (src/ORBmatcher.cc:1218-1350) and
ORBmatcher::SearchByProjection(F, vpMapPoints, th, checkLevel) (:43-126) — assignment arrays and counts identical, including the
order-dependent cases (queries competing for the same feature), which the kernel resolves in reservation rounds."""
import ctypes as C
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
K = (458.654, 457.296, 367.215, 248.375)
W, Hh = 752, 480


def _oracle_last(gpu, oracle, ka, da, kb, db, wp, valid, pose_l, pose_c, ur, th, mono, check_level, check_ori, mb, mbf):
    mp_last = np.where(valid > 0, np.arange(len(ka)), -1).astype(np.int32)
    cur = gpu.make_frame_view(kb, db, u_right=ur, K=K, mb=mb, mbf=mbf, scale_factors=H.SCALE8, Rcw=pose_c[:9].reshape(3, 3), tcw=pose_c[9:],
                              cls=oracle.FrameView)
    last = gpu.make_frame_view(ka, da, map_point=mp_last, K=K, scale_factors=H.SCALE8, Rcw=pose_l[:9].reshape(3, 3), tcw=pose_l[9:],
                               cls=oracle.FrameView)
    mps = gpu.make_map_points(wp[:len(ka)], da, cls=oracle.MapPoints)
    n = oracle.lib().orc_search_by_projection_last(int(check_ori), cur.ptr(), last.ptr(), mps.ptr(), th, int(mono), int(check_level))
    return n, cur.map_point.copy()


@pytest.mark.parametrize("th,mono,check_level,check_ori,tz", [(15.0, True, True, True, 0.0), (7.0, False, True, True, 0.4), (7.0, False, True, False, -0.4),
                                                              (30.0, True, False, True, 0.0)])
def test_search_by_projection_last_batch(gpu, oracle, synth, th, mono, check_level, check_ori, tz):
    B = 5
    imgs = synth.synth_stream(21, B, W, Hh)
    ex = gpu.ORBextractor(1000, 1.2, 8, 20, 7, width=W, height=Hh, max_batch=B)
    res = ex.extract_batch(imgs)
    cap = ex.cap
    rng = np.random.default_rng(5)
    fx, fy, cx, cy = K
    wp = np.zeros((B, cap, 3), np.float32); valid = np.zeros((B, cap), np.uint8)
    pose = np.tile(np.concatenate([np.eye(3, dtype=np.float32).reshape(-1), np.zeros(3, np.float32)]), (B, 1)).astype(np.float32)
    pose[:, 11] = np.arange(B) * tz                    # camera advancing / receding along z: forward / backward level bands when stereo
    ur = None
    if not mono:
        ur = np.full((B, cap), -1, np.float32)
    for f, (k, d) in enumerate(res):
        n = len(k)
        z = rng.uniform(2, 10, n)
        u = k["x"].astype(np.float64) + rng.normal(0, 0.7, n); v = k["y"].astype(np.float64) + rng.normal(0, 0.7, n)
        # world point = back-projection through the frame's own pose (R = I, t = (0, 0, tz f))
        wp[f, :n] = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z - pose[f, 11]], -1)
        valid[f, :n] = rng.random(n) < 0.8
        if ur is not None:
            ur[f, :n] = np.where(rng.random(n) < 0.6, k["x"] - 40.0 / rng.uniform(2, 10, n), -1)
    last = np.array([0, 1, 2, 3, 0], np.int32); cur = np.array([1, 2, 3, 4, 4], np.int32)
    mb, mbf = 0.09, 40.0
    m = gpu.ORBmatcher(0.9, check_ori)
    nm, out, rounds = m.SearchByProjectionLastBatch(ex, last, cur, wp, valid, pose, K=K, mb=mb, mbf=mbf, u_right=ur, th=th, mono=mono,
                                                   check_level=check_level, check_ori=check_ori)
    for p, (a, b) in enumerate(zip(last, cur)):
        (ka, da), (kb, db) = res[a], res[b]
        on, omap = _oracle_last(gpu, oracle, ka, da, kb, db, wp[a], valid[a][:len(ka)], pose[a], pose[b], None if ur is None else ur[b][:len(kb)],
                                th, mono, check_level, check_ori, mb, mbf)
        assert nm[p] == on, (p, nm[p], on, rounds[p])
        assert (out[p, :len(kb)] == omap).all(), (p, int((out[p, :len(kb)] != omap).sum()))
        assert (out[p, len(kb):] == -1).all()
    assert nm[:4].min() > 100 and rounds.max() >= 2      # competing queries did occur and were resolved (pair 4 spans four frames)


def test_search_by_projection_last_batch_crowded(gpu, oracle, synth):
    """Every last-frame point projects into the same small region: long chains of queries competing for the same features."""
    B = 2
    imgs = synth.synth_stream(23, B, W, Hh)
    ex = gpu.ORBextractor(1000, 1.2, 8, 20, 7, width=W, height=Hh, max_batch=B)
    res = ex.extract_batch(imgs)
    cap = ex.cap
    rng = np.random.default_rng(6)
    fx, fy, cx, cy = K
    (ka, da), (kb, db) = res
    n = len(ka)
    wp = np.zeros((B, cap, 3), np.float32); valid = np.zeros((B, cap), np.uint8)
    u = 300 + rng.uniform(-20, 20, n); v = 240 + rng.uniform(-20, 20, n)
    wp[0, :n] = np.stack([(u - cx) * 5 / fx, (v - cy) * 5 / fy, np.full(n, 5.0)], -1)
    valid[0, :n] = 1
    pose = np.tile(np.concatenate([np.eye(3, dtype=np.float32).reshape(-1), np.zeros(3, np.float32)]), (B, 1)).astype(np.float32)
    m = gpu.ORBmatcher(0.9, True)
    nm, out, rounds = m.SearchByProjectionLastBatch(ex, [0], [1], wp, valid, pose, K=K, th=15.0, mono=True, check_level=False, check_ori=False)
    on, omap = _oracle_last(gpu, oracle, ka, da, kb, db, wp[0], valid[0][:n], pose[0], pose[1], None, 15.0, True, False, False, 0.0, 0.0)
    assert nm[0] == on and (out[0, :len(kb)] == omap).all()
    assert rounds[0] > 3 and on > 20


@pytest.mark.parametrize("th,check_level,stereo", [(1.0, True, False), (3.0, False, False), (5.0, True, True)])
def test_search_by_projection_points_batch(gpu, oracle, synth, th, check_level, stereo):
    B = 4
    imgs = synth.synth_stream(25, B + 1, W, Hh)
    ex = gpu.ORBextractor(1000, 1.2, 8, 20, 7, width=W, height=Hh, max_batch=B + 1)
    res = ex.extract_batch(imgs)
    cap = ex.cap
    rng = np.random.default_rng(7)
    M = 900
    frame = np.arange(1, B + 1, dtype=np.int32)
    px = np.zeros((B, M), np.float32); py = np.zeros((B, M), np.float32); pxr = np.zeros((B, M), np.float32)
    vc = np.zeros((B, M), np.float32); lv = np.zeros((B, M), np.int32); us = np.zeros((B, M), np.uint8)
    md = np.zeros((B, M, 32), np.uint8); occ = np.zeros((B, cap), np.uint8)
    ur = np.full((B + 1, cap), -1, np.float32) if stereo else None
    for p, f in enumerate(frame):
        ka, da = res[f - 1]                               # the map points: the previous frame's features seen again
        kb, db = res[f]
        m = min(M, len(ka))
        px[p, :m] = ka["x"][:m] + rng.normal(0, 1.0, m); py[p, :m] = ka["y"][:m] + rng.normal(0, 1.0, m)
        pxr[p, :m] = px[p, :m] - 8.0
        vc[p, :m] = rng.uniform(0.99, 1.0, m); lv[p, :m] = ka["octave"][:m]; us[p, :m] = rng.random(m) < 0.9
        md[p, :m] = da[:m]
        occ[p, :len(kb):9] = 1
        if stereo:
            ur[f, :len(kb)] = np.where(rng.random(len(kb)) < 0.6, kb["x"] - 8.0 + rng.normal(0, 2.0, len(kb)), -1)
    mt = gpu.ORBmatcher(0.8)
    nm, out, rounds = mt.SearchByProjectionPointsBatch(ex, frame, px, py, pxr, vc, lv, us, md, occupied=occ, u_right=ur, K=K, mb=0.09, mbf=40.0,
                                                       nnratio=0.8, th=th, check_level=check_level)
    for p, f in enumerate(frame):
        kb, db = res[f]
        pre = np.where(occ[p, :len(kb)] > 0, M, -1).astype(np.int32)      # pre-occupied slots: an extra map point (index M) with observations
        track = dict(in_view=np.r_[us[p], 0].astype(np.uint8), x=np.r_[px[p], 0].astype(np.float32), y=np.r_[py[p], 0].astype(np.float32),
                     xr=np.r_[pxr[p], 0].astype(np.float32), level=np.r_[lv[p], 0].astype(np.int32), view_cos=np.r_[vc[p], 0].astype(np.float32))
        fv = gpu.make_frame_view(kb, db, map_point=pre.copy(), u_right=None if ur is None else ur[f][:len(kb)], K=K, mbf=40.0, mb=0.09,
                                 scale_factors=H.SCALE8, cls=oracle.FrameView)
        mps = gpu.make_map_points(np.zeros((M + 1, 3)), np.concatenate([md[p], np.zeros((1, 32), np.uint8)]), observations=np.ones(M + 1, np.int32),
                                  track=track, cls=oracle.MapPoints)
        q = np.arange(M, dtype=np.int32)
        on = oracle.lib().orc_search_by_projection_points(0.8, fv.ptr(), mps.ptr(), q.ctypes.data_as(oracle.i32p), M, th, int(check_level))
        want = fv.map_point.copy()
        want[pre >= 0] = -1                                # the batched form reports new assignments only
        assert nm[p] == on, (p, nm[p], on)
        assert (out[p, :len(kb)] == want).all(), (p, int((out[p, :len(kb)] != want).sum()))
    assert nm.min() > 50
