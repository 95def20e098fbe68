"""GPU parity of the ORBmatcher Hamming search (through the C ABI) against the oracle: bit-exact
distances, best/second-best, match index arrays and match counts for every Search* variant."""
import numpy as np
import pytest
import helpers as H

pytestmark = pytest.mark.gpu
K = (458.654, 457.296, 367.215, 248.375)


@pytest.fixture(scope="module")
def frames(oracle, synth):
    return H.two_frames(oracle, synth, seed=1)


def test_bf_top2_and_pairs(gpu, oracle, synth):
    m = gpu.ORBmatcher()
    for nq, nt in [(1000, 1000), (2000, 2000), (37, 5), (1, 1), (300, 0)]:
        q, t, _ = synth.synth_descriptors(nq + nt, nq, max(nt, 1))
        t = t[:nt]
        if nt > 3:
            t[3] = t[1]                      # exact duplicates: first index must win, second == best
            q[0] = t[1]
        bi, bd, sd = m.bf_top2(q, t)
        if nt == 0:
            assert (bi == -1).all() and (bd == 256).all()
            continue
        obi, obd, osd = oracle.bf_top2(q, t)
        assert (bi == obi).all() and (bd == obd).all() and (sd == osd).all()
    q, t, _ = synth.synth_descriptors(3, 500, 700)
    rng = np.random.default_rng(0)
    pq = rng.integers(0, 500, 20000).astype(np.int32)
    pt = rng.integers(0, 700, 20000).astype(np.int32)
    d = m.pairs(q, t, pq, pt)
    ref = np.unpackbits(q[pq] ^ t[pt], axis=1).sum(1)
    assert (d == ref).all()
    assert gpu.ORBmatcher.DescriptorDistance(q[0], t[0]) == int(np.unpackbits(q[0] ^ t[0]).sum())


def _views(mod, oracle_cls, ka, da, kb, db, mp_a=None, **kw):
    """(kf view, f view) for module `mod` (package or oracle_lib struct classes)."""
    kf = mod.make_frame_view(ka, da, map_point=mp_a, node_of_feature=H.synth_vocab_nodes(da), cls=oracle_cls[0], **kw)
    f = mod.make_frame_view(kb, db, node_of_feature=H.synth_vocab_nodes(db), cls=oracle_cls[0], **kw)
    return kf, f


@pytest.mark.parametrize("nnratio,check_ori", [(0.7, False), (0.75, True), (0.9, True)])
def test_search_by_bow(gpu, oracle, frames, nnratio, check_ori):
    (_, ka, da), (_, kb, db) = frames
    rng = np.random.default_rng(2)
    mp = np.arange(len(ka), dtype=np.int32)
    mp[rng.random(len(ka)) < 0.2] = -1                     # KF features without a map point
    bad = (rng.random(len(ka)) < 0.05).astype(np.uint8)    # isBad()
    m = gpu.ORBmatcher(nnratio, check_ori)
    kf, f = _views(gpu, (gpu.FrameView,), ka, da, kb, db, mp_a=mp)
    mps = gpu.make_map_points(np.zeros((len(ka), 3)), da, bad=bad)
    n, out = m.SearchByBoW(kf, mps, f)
    okf, of = _views(gpu, (oracle.FrameView,), ka, da, kb, db, mp_a=mp)
    omps = gpu.make_map_points(np.zeros((len(ka), 3)), da, bad=bad, cls=oracle.MapPoints)
    oout = np.zeros(len(kb), np.int32)
    on = oracle.lib().orc_search_by_bow(nnratio, int(check_ori), okf.ptr(), omps.ptr(), of.ptr(),
                                        oout.ctypes.data_as(oracle.i32p))
    assert n == on and (out == oout).all()
    assert n > 50          # the scene overlap must actually produce matches


def _projection_setup(ka, da, kb, shift, rng):
    """Map points = KF features back-projected at random depth so that they project to kp - shift in the
    current frame (identity pose)."""
    fx, fy, cx, cy = K
    m = len(ka)
    z = rng.uniform(2, 10, m)
    u = ka["x"].astype(np.float64) - shift[0] + rng.normal(0, 1.0, m)
    v = ka["y"].astype(np.float64) - shift[1] + rng.normal(0, 1.0, m)
    wp = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], -1).astype(np.float32)
    track = dict(in_view=(rng.random(m) < 0.9).astype(np.uint8), x=u.astype(np.float32), y=v.astype(np.float32),
                 xr=(u - 40.0 / z).astype(np.float32), level=ka["octave"].astype(np.int32),
                 view_cos=rng.uniform(0.99, 1.0, m).astype(np.float32))
    obs = (rng.random(m) < 0.85).astype(np.int32) * 3
    return wp, track, obs, z


@pytest.mark.parametrize("th,check_level,stereo", [(1.0, True, False), (3.0, False, False), (5.0, True, True)])
def test_search_by_projection_points(gpu, oracle, frames, th, check_level, stereo):
    (_, ka, da), (_, kb, db) = frames
    rng = np.random.default_rng(3)
    wp, track, obs, z = _projection_setup(ka, da, kb, (9, 5), rng)
    ur = (kb["x"] - 30.0).astype(np.float32) if stereo else None
    if stereo:
        ur[::3] = -1
    pre = np.full(len(kb), -1, np.int32)
    pre[::11] = rng.integers(0, len(ka), len(pre[::11]))       # some features already hold a map point
    q = rng.permutation(len(ka)).astype(np.int32)
    res = []
    for mod_cls, mp_cls, run in (((gpu.FrameView,), gpu.MapPoints, "gpu"), ((oracle.FrameView,), oracle.MapPoints, "orc")):
        f = gpu.make_frame_view(kb, db, map_point=pre.copy(), u_right=ur, K=K, mbf=40.0, mb=0.09, scale_factors=H.SCALE8,
                                cls=mod_cls[0])
        mps = gpu.make_map_points(wp, da, observations=obs, track=track, cls=mp_cls)
        if run == "gpu":
            n = gpu.ORBmatcher(0.8).SearchByProjectionPoints(f, mps, q, th, check_level)
        else:
            n = oracle.lib().orc_search_by_projection_points(0.8, f.ptr(), mps.ptr(), q.ctypes.data_as(oracle.i32p), len(q),
                                                             th, int(check_level))
        res.append((n, f.map_point.copy()))
    assert res[0][0] == res[1][0] and (res[0][1] == res[1][1]).all()
    assert res[0][0] > 20


@pytest.mark.parametrize("mono,check_level,check_ori", [(True, True, True), (False, True, True), (True, False, False)])
def test_search_by_projection_last_and_kf(gpu, oracle, frames, mono, check_level, check_ori):
    (_, ka, da), (_, kb, db) = frames
    rng = np.random.default_rng(4)
    wp, track, obs, z = _projection_setup(ka, da, kb, (9, 5), rng)
    mp_last = np.arange(len(ka), dtype=np.int32)
    mp_last[rng.random(len(ka)) < 0.3] = -1
    outl = (rng.random(len(ka)) < 0.05).astype(np.uint8)
    tl = np.array([0.0, 0.0, 0.3 if not mono else 0.0], np.float32)     # last-frame pose: forward motion when stereo
    found = (rng.random(len(ka)) < 0.1).astype(np.uint8)
    maxd = (z * 1.5).astype(np.float32)
    mind = (z * 0.5).astype(np.float32)
    out = {}
    for name, fcls, mcls in (("gpu", gpu.FrameView, gpu.MapPoints), ("orc", oracle.FrameView, oracle.MapPoints)):
        cur = gpu.make_frame_view(kb, db, K=K, mb=0.09, mbf=40.0, scale_factors=H.SCALE8, cls=fcls)
        last = gpu.make_frame_view(ka, da, map_point=mp_last, outlier=outl, K=K, tcw=tl, scale_factors=H.SCALE8, cls=fcls)
        mps = gpu.make_map_points(wp, da, observations=obs, min_dist=mind, max_dist=maxd, cls=mcls)
        cur2 = gpu.make_frame_view(kb, db, K=K, scale_factors=H.SCALE8, cls=fcls)
        kfv = gpu.make_frame_view(ka, da, map_point=mp_last, K=K, scale_factors=H.SCALE8, cls=fcls)
        if name == "gpu":
            m = gpu.ORBmatcher(0.9, check_ori)
            n1 = m.SearchByProjectionLast(cur, last, mps, 15.0, mono, check_level)
            n2 = m.SearchByProjectionKF(cur2, kfv, mps, found, 10.0, 100)
        else:
            L = oracle.lib()
            n1 = L.orc_search_by_projection_last(int(check_ori), cur.ptr(), last.ptr(), mps.ptr(), 15.0, int(mono), int(check_level))
            n2 = L.orc_search_by_projection_kf(int(check_ori), cur2.ptr(), kfv.ptr(), mps.ptr(), found.ctypes.data_as(oracle.u8p), 10.0, 100)
        out[name] = (n1, cur.map_point.copy(), n2, cur2.map_point.copy())
    g, o = out["gpu"], out["orc"]
    assert g[0] == o[0] and (g[1] == o[1]).all()
    assert g[2] == o[2] and (g[3] == o[3]).all()
    assert g[0] > 20 and g[2] > 20


def test_search_for_initialization(gpu, oracle, frames):
    (_, ka, da), (_, kb, db) = frames
    prev = np.stack([ka["x"], ka["y"]], -1).astype(np.float32)
    f1g = gpu.make_frame_view(ka, da, K=K, scale_factors=H.SCALE8)
    f2g = gpu.make_frame_view(kb, db, K=K, scale_factors=H.SCALE8)
    n, m12, prev_out = gpu.ORBmatcher(0.9, True).SearchForInitialization(f1g, f2g, prev, 100)
    f1o = gpu.make_frame_view(ka, da, K=K, scale_factors=H.SCALE8, cls=oracle.FrameView)
    f2o = gpu.make_frame_view(kb, db, K=K, scale_factors=H.SCALE8, cls=oracle.FrameView)
    oprev = prev.copy()
    om12 = np.zeros(len(ka), np.int32)
    on = oracle.lib().orc_search_for_initialization(0.9, 1, f1o.ptr(), f2o.ptr(), oprev.ctypes.data_as(oracle.f32p),
                                                    om12.ctypes.data_as(oracle.i32p), 100)
    assert n == on and (m12 == om12).all() and (prev_out == oprev).all()
    assert n > 20


def test_search_by_bow_batch_device(gpu, oracle, synth):
    """Extractor batch -> device-resident SearchByBoW over consecutive frame pairs, vs the oracle per pair."""
    w, h = 752, 480
    imgs = synth.synth_stream(11, 4, w, h)
    ex = gpu.ORBextractor(1000, 1.2, 8, 20, 7, width=w, height=h, max_batch=4)
    res = ex.extract_batch(imgs)
    cap = ex.cap
    nof = np.full((4, cap), 0xFFFFFFFF, np.uint32)
    for f, (k, d) in enumerate(res):
        nof[f, :len(k)] = H.synth_vocab_nodes(d)
    m = gpu.ORBmatcher(0.7, True)
    kf_idx, f_idx = [0, 1, 2, 0], [1, 2, 3, 3]
    nm, out = m.SearchByBoWBatch(ex, kf_idx, f_idx, nof)
    for p, (a, b) in enumerate(zip(kf_idx, f_idx)):
        ka, da = res[a]
        kb, db = res[b]
        okf = gpu.make_frame_view(ka, da, map_point=np.arange(len(ka), dtype=np.int32), node_of_feature=nof[a], cls=oracle.FrameView)
        of = gpu.make_frame_view(kb, db, node_of_feature=nof[b], cls=oracle.FrameView)
        omps = gpu.make_map_points(np.zeros((len(ka), 3)), da, cls=oracle.MapPoints)
        oout = np.zeros(len(kb), np.int32)
        on = oracle.lib().orc_search_by_bow(0.7, 1, okf.ptr(), omps.ptr(), of.ptr(), oout.ctypes.data_as(oracle.i32p))
        assert nm[p] == on, (p, nm[p], on)
        assert (out[p, :len(kb)] == oout).all()
        assert (out[p, len(kb):] == -1).all()
    assert nm.min() > 30


def test_search_by_bow_batch_synthetic_vocabulary(gpu, oracle, synth):
    """node_of_feature=None -> node = descriptor byte0 >> 2 (bench stand-in); same semantics, checked vs the oracle."""
    w, h = 752, 480
    imgs = synth.synth_stream(13, 3, w, h)
    ex = gpu.ORBextractor(1000, 1.2, 8, 20, 7, width=w, height=h, max_batch=3)
    res = ex.extract_batch(imgs)
    m = gpu.ORBmatcher(0.7, True)
    kf_idx, f_idx = [2, 0, 1], [0, 1, 2]
    nm, out = m.SearchByBoWBatch(ex, kf_idx, f_idx, None)
    for p, (a, b) in enumerate(zip(kf_idx, f_idx)):
        ka, da = res[a]
        kb, db = res[b]
        okf = gpu.make_frame_view(ka, da, map_point=np.arange(len(ka), dtype=np.int32), node_of_feature=(da[:, 0] >> 2).astype(np.uint32),
                                  cls=oracle.FrameView)
        of = gpu.make_frame_view(kb, db, node_of_feature=(db[:, 0] >> 2).astype(np.uint32), cls=oracle.FrameView)
        omps = gpu.make_map_points(np.zeros((len(ka), 3)), da, cls=oracle.MapPoints)
        oout = np.zeros(len(kb), np.int32)
        on = oracle.lib().orc_search_by_bow(0.7, 1, okf.ptr(), omps.ptr(), of.ptr(), oout.ctypes.data_as(oracle.i32p))
        assert nm[p] == on and (out[p, :len(kb)] == oout).all()
    assert nm.min() > 30
