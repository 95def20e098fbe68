"""Scene for ORBmatcher::Fuse(KeyFrame*, vector<MapPoint*>, th) (src/ORBmatcher.cc:748-886): a (possibly stereo) key frame and candidate
map points of a neighbouring key frame that project near its key points.  Shared by the CPU pin test and the GPU tests."""
import numpy as np

K = (458.654, 457.296, 367.215, 248.375)
BOUNDS = (0.0, 0.0, 752.0, 480.0)
MBF = 47.9


def rot(ax, ang):
    ax = np.asarray(ax, np.float64); ax /= np.linalg.norm(ax)
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx


def camera_center(R32, t32):
    """KeyFrame::SetPose (src/KeyFrame.cc:100-119): Ow = -(Rwc * tcw), float, accumulated left to right."""
    f = np.float32
    Ow = np.zeros(3, np.float32)
    for i in range(3):
        a = f(R32[0, i]) * f(t32[0])
        a = f(a + f(R32[1, i]) * f(t32[1]))
        a = f(a + f(R32[2, i]) * f(t32[2]))
        Ow[i] = -a
    return Ow


def build(ka, da, kb, db, seed, stereo):
    """Points come from frame a's keys (descriptor = da), shifted so that they land on frame b's keys (the key frame).  Returns a dict of numpy
    arrays; `ids` is the vpMapPoints order with NULL entries (-1) and duplicates."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = K
    R = rot((0.2, 1.0, 0.1), 0.03); t = np.array([0.05, -0.02, 0.1])
    R32, t32 = R.astype(np.float32), t.astype(np.float32)
    m = len(ka)
    z = rng.uniform(2, 10, m)
    u = ka["x"].astype(np.float64) - 9.0 + rng.normal(0, 0.8, m); v = ka["y"].astype(np.float64) - 5.0 + rng.normal(0, 0.8, m)
    Xc = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], -1)
    Xw = (Xc - t) @ R
    Ow64 = -R.T @ t
    PO = Xw - Ow64
    dist = np.linalg.norm(PO, axis=1)
    normal = PO / dist[:, None]
    normal[rng.random(m) < 0.05] *= -1                      # viewing-angle rejections
    maxd = dist * 1.2 ** (ka["octave"] + 0.5)
    mind = dist * 0.4
    far = rng.random(m) < 0.05
    mind[far] = dist[far] * 2.0                             # depth outside the invariance region
    behind = rng.random(m) < 0.02
    Xw[behind] = (np.array([0.0, 0.0, -3.0]) - t) @ R       # negative depth
    nb = len(kb)
    u_right = None
    if stereo:                                              # right coordinate consistent with SOME depth in range (+ noise -> chi2 rejections)
        zr = rng.uniform(2, 10, nb)
        u_right = (kb["x"] - MBF / zr + rng.normal(0, 0.6, nb)).astype(np.float32)
        u_right[rng.random(nb) < 0.4] = -1.0                # monocular keys
    extra = 60                                              # map points already sitting in key-frame slots
    Xw2 = np.concatenate([Xw, rng.uniform(-1, 1, (extra, 3))]).astype(np.float32)
    nrm2 = np.concatenate([normal, np.tile([0, 0, 1.0], (extra, 1))]).astype(np.float32)
    desc = np.concatenate([da, rng.integers(0, 256, (extra, 32), dtype=np.uint8)])
    bad = (rng.random(m + extra) < 0.05).astype(np.uint8)
    slots = np.full(nb, -1, np.int32)
    occ = rng.permutation(nb)[:extra]
    slots[occ] = m + np.arange(extra)                       # each extra point in exactly one slot
    some = rng.permutation(m)[:30]                          # a few candidates are already observed by the key frame (IsInKeyFrame skip)
    free = np.setdiff1d(np.arange(nb), occ)[:30]
    slots[free] = some
    obs2 = np.full(m + extra, -1, np.int32)                 # observations in the second key frame: makes Observations() differ
    have2 = rng.permutation(m + extra)[: (m + extra) * 2 // 3]
    obs2[have2] = np.arange(len(have2))
    ids = rng.permutation(m).astype(np.int32)
    ids = np.concatenate([ids, ids[:25]])                   # duplicates: second visit must see the first one's effect
    ids[rng.permutation(len(ids))[:20]] = -1                # NULL entries
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    inv_sigma2 = (np.float32(1.0) / (sf * sf)).astype(np.float32)
    return dict(R=R32, t=t32, Ow=camera_center(R32, t32), Xw=Xw2, normal=nrm2, desc=desc, bad=bad, mind=np.r_[mind, np.ones(extra)].astype(np.float32),
                maxd=np.r_[maxd, np.ones(extra)].astype(np.float32), slots=slots, obs2=obs2, n2=int(len(have2)), ids=ids, u_right=u_right,
                inv_sigma2=inv_sigma2, sf=sf, m=m + extra)


def run_full(fn, pkg, oracle, kb, db, S, th):
    """refm_fuse_kf / adp_fuse_kf on scene S -> (nFused, key-frame slots, bad, n_obs, replaced_by, slots of the second key frame)."""
    import ctypes as C
    mps = pkg.make_map_points(S["Xw"], S["desc"], bad=S["bad"], min_dist=S["mind"], max_dist=S["maxd"], normal=S["normal"], cls=oracle.MapPoints)
    kf = pkg.make_frame_view(kb, db, map_point=S["slots"].copy(), u_right=S["u_right"], K=K, mbf=MBF, bounds=BOUNDS, scale_factors=S["sf"],
                             Rcw=S["R"], tcw=S["t"], cls=oracle.FrameView)
    m = S["m"]
    bad = np.zeros(m, np.uint8); nobs = np.zeros(m, np.int32); rep = np.zeros(m, np.int32); k2 = np.zeros(max(S["n2"], 1), np.int32)
    f32p, i32p, u8p = oracle.f32p, oracle.i32p, C.POINTER(C.c_uint8)
    ids = S["ids"]
    fn.restype = C.c_int
    n = fn(C.cast(kf.ptr(), C.c_void_p), S["inv_sigma2"].ctypes.data_as(f32p), C.cast(mps.ptr(), C.c_void_p), S["obs2"].ctypes.data_as(i32p), S["n2"],
           ids.ctypes.data_as(i32p), len(ids), C.c_float(th), bad.ctypes.data_as(u8p), nobs.ctypes.data_as(i32p), rep.ctypes.data_as(i32p),
           k2.ctypes.data_as(i32p))
    return n, kf.map_point.copy(), bad, nobs, rep, k2
