"""ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:748-886, local mapping).  The function is a per-point
search (projection, gates, best descriptor in a window) followed by a mutation of the observation graph.  The product computes the search
on the device (ygz_fuse_candidates) and the adapter replays the mutation on the caller's objects; here the restatement of the search
(orc_fuse_candidates) is pinned to the REFERENCE's own compiled code (oracle/_ref/libmatcher_ref.so, one point at a time against an empty
key frame: the index Fuse adds an observation at).  Build container only."""
import ctypes as C
import numpy as np
import pytest

import fuse_scene as FS
import helpers as H
import ref_lib


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


@pytest.mark.parametrize("stereo,th,seed", [(False, 3.0, 21), (True, 3.0, 22), (True, 6.0, 23), (False, 1.5, 24)])
def test_fuse_candidates_pinned_to_reference(ref, pkg, oracle, frames, stereo, th, seed):
    (_, ka, da), (_, kb, db) = frames
    S = FS.build(ka, da, kb, db, seed, stereo)
    mps = pkg.make_map_points(S["Xw"], S["desc"], bad=S["bad"], min_dist=S["mind"], max_dist=S["maxd"], normal=S["normal"], cls=oracle.MapPoints)
    kf = pkg.make_frame_view(kb, db, u_right=S["u_right"], K=FS.K, mbf=FS.MBF, bounds=FS.BOUNDS, scale_factors=S["sf"], Rcw=S["R"], tcw=S["t"],
                             cls=oracle.FrameView)
    ids = S["ids"]
    n = len(ids)
    bi = np.zeros(n, np.int32); bd = np.zeros(n, np.int32); rbi = np.zeros(n, np.int32)
    f32p, i32p = oracle.f32p, oracle.i32p
    oracle.lib().orc_fuse_candidates(vp(kf), S["Ow"].ctypes.data_as(f32p), S["inv_sigma2"].ctypes.data_as(f32p), vp(mps), ids.ctypes.data_as(i32p), n,
                                     C.c_float(th), bi.ctypes.data_as(i32p), bd.ctypes.data_as(i32p))
    ref.refm_fuse_kf_each(vp(kf), S["inv_sigma2"].ctypes.data_as(f32p), vp(mps), ids.ctypes.data_as(i32p), n, C.c_float(th), rbi.ctypes.data_as(i32p))
    mine = np.where(bd <= 50, bi, -1)                       # ORBmatcher::TH_LOW
    assert (mine == rbi).all(), np.nonzero(mine != rbi)[0][:10]
    assert (rbi >= 0).sum() > 200 and (rbi < 0).sum() > 40
    assert ((bi >= 0) & (bd > 50)).any()                    # candidates found but too far in descriptor space
    if stereo:                                              # the stereo chi2 gate changes decisions
        kfm = pkg.make_frame_view(kb, db, K=FS.K, mbf=FS.MBF, bounds=FS.BOUNDS, scale_factors=S["sf"], Rcw=S["R"], tcw=S["t"], cls=oracle.FrameView)
        bim = np.zeros(n, np.int32); bdm = np.zeros(n, np.int32)
        oracle.lib().orc_fuse_candidates(vp(kfm), S["Ow"].ctypes.data_as(f32p), S["inv_sigma2"].ctypes.data_as(f32p), vp(mps), ids.ctypes.data_as(i32p), n,
                                         C.c_float(th), bim.ctypes.data_as(i32p), bdm.ctypes.data_as(i32p))
        assert (bim != bi).any()


@pytest.mark.parametrize("stereo,th,seed", [(False, 3.0, 21), (True, 3.0, 22)])
def test_reference_fuse_exercises_every_branch(ref, pkg, oracle, frames, stereo, th, seed):
    """Characterises the scene the GPU adapter test replays (tests/test_fuse_gpu.py compares the adapter's final map state with this one's):
    the reference's own Fuse must take every branch of :866-880 on it."""
    (_, ka, da), (_, kb, db) = frames
    S = FS.build(ka, da, kb, db, seed, stereo)
    n, slots, bad, nobs, rep, k2 = FS.run_full(ref.refm_fuse_kf, pkg, oracle, kb, db, S, th)
    assert n > 150
    newly_bad = np.nonzero((bad != 0) & (S["bad"] == 0))[0]
    assert len(newly_bad) > 5 and (rep[newly_bad] >= 0).all()
    m0 = len(ka)
    assert (newly_bad < m0).any() and (newly_bad >= m0).any()          # pMP->Replace(pMPinKF) and pMPinKF->Replace(pMP)
    filled = (S["slots"] < 0) & (slots >= 0)
    assert filled.sum() > 50                                           # AddObservation + AddMapPoint
    assert nobs.max() >= 3 if stereo else nobs.max() >= 2


def _replay(S, nb, bi, bd):
    """The adapter's half of Fuse (include/ygz/adapters.h ORBmatcherT::Fuse) on a plain-Python copy of the shim's two-key-frame map: the loop of
    :762-884 with the per-point search replaced by a table computed BEFORE the loop (bi / bd), skip tests and map mutation live."""
    m = S["m"]
    slots = [S["slots"].copy(), np.full(max(S["n2"], 1), -1, np.int32)]
    ur = S["u_right"]
    obs = [dict() for _ in range(m)]
    nobs = np.zeros(m, np.int32); bad = S["bad"].copy(); rep = np.full(m, -1, np.int32)

    def add_obs(j, k, idx):
        if k in obs[j]:
            return
        obs[j][k] = idx
        nobs[j] += 2 if (k == 0 and ur is not None and ur[idx] >= 0) else 1

    def replace(a, b):                                   # a->Replace(b), src/MapPoint.cc:155-188
        if a == b:
            return
        o, obs[a] = obs[a], {}
        bad[a] = 1; rep[a] = b
        for k, idx in o.items():
            if k not in obs[b]:
                slots[k][idx] = b; add_obs(b, k, idx)
            else:
                slots[k][idx] = -1

    for i in range(nb):
        if slots[0][i] >= 0:
            add_obs(slots[0][i], 0, i)
    for j in range(m):
        if S["obs2"][j] >= 0:
            slots[1][S["obs2"][j]] = j; add_obs(j, 1, int(S["obs2"][j]))
    n = 0
    for i, j in enumerate(S["ids"]):
        if j < 0 or bad[j] or 0 in obs[j]:
            continue
        if bd[i] <= 50:
            idx = int(bi[i]); inkf = int(slots[0][idx])
            if inkf >= 0:
                if not bad[inkf]:
                    if nobs[inkf] > nobs[j]:
                        replace(j, inkf)
                    else:
                        replace(inkf, j)
            else:
                add_obs(j, 0, idx); slots[0][idx] = j
            n += 1
    return n, slots[0], bad, nobs, rep, slots[1]


@pytest.mark.parametrize("stereo,th,seed", [(False, 3.0, 21), (True, 3.0, 22), (True, 5.0, 26), (False, 1.5, 27)])
def test_search_ahead_of_the_mutations_equals_the_interleaved_loop(ref, pkg, oracle, frames, stereo, th, seed):
    """The product computes every point's best key point BEFORE the loop that mutates the map (DESIGN.md §7 argues why that is the same).
    Here: the restatement's one-shot search + a replay of the loop (the adapter's logic, in Python) against the REFERENCE's own compiled
    Fuse, which searches and mutates interleaved — same return value and same final map (slots of both key frames, bad flags, observation
    counts, mpReplaced), with duplicates, NULL entries, points already in the key frame and replacements in both directions in play."""
    (_, ka, da), (_, kb, db) = frames
    S = FS.build(ka, da, kb, db, seed, stereo)
    mps = pkg.make_map_points(S["Xw"], S["desc"], bad=S["bad"], min_dist=S["mind"], max_dist=S["maxd"], normal=S["normal"], cls=oracle.MapPoints)
    kf = pkg.make_frame_view(kb, db, u_right=S["u_right"], K=FS.K, mbf=FS.MBF, bounds=FS.BOUNDS, scale_factors=S["sf"], Rcw=S["R"], tcw=S["t"],
                             cls=oracle.FrameView)
    ids = S["ids"]; n = len(ids)
    bi = np.zeros(n, np.int32); bd = np.zeros(n, np.int32)
    f32p, i32p = oracle.f32p, oracle.i32p
    oracle.lib().orc_fuse_candidates(vp(kf), S["Ow"].ctypes.data_as(f32p), S["inv_sigma2"].ctypes.data_as(f32p), vp(mps), ids.ctypes.data_as(i32p), n,
                                     C.c_float(th), bi.ctypes.data_as(i32p), bd.ctypes.data_as(i32p))
    mine = _replay(S, len(kb), bi, bd)
    theirs = FS.run_full(ref.refm_fuse_kf, pkg, oracle, kb, db, S, th)
    assert mine[0] == theirs[0] and mine[0] > 100
    for x, y, what in zip(mine[1:], theirs[1:], ("key-frame slots", "bad", "observations", "replaced_by", "second key frame")):
        assert (np.asarray(x) == np.asarray(y)).all(), what
