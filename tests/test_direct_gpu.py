"""GPU parity of the batched ORBmatcher::FindDirectProjection (src/ORBmatcher.cc:1575-1602: GetWarpAffineMatrix,
GetBestSearchLevel, WarpAffine, ygz::Align2D) against the oracle: refined pixel bit-exact (the device keeps the
reference's sequential float accumulation), search level and return value identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(gpu, synth, seed, w, h, nlevels, scale, n_items, **kw):
    s = synth.synth_direct_scene(seed, n_items=n_items, w=w, h=h, nlevels=nlevels, scale=scale, **kw)
    ex = gpu.ORBextractor(500, scale, nlevels, 20, 7, width=w, height=h)
    t = ex.tables()

    def pyr(img):
        ex.ComputePyramid(img)
        return [ex.level(0, l) for l in range(nlevels)]
    kfs = [gpu.make_direct_frame(pyr(im), q, tt) for im, (q, tt) in zip(s["kf_imgs"], s["kf_poses"])]
    cur = gpu.make_direct_frame(pyr(s["cur_img"]), *s["cur_pose"])
    return s, t, kfs, cur


@pytest.mark.parametrize("seed,w,h,nlevels,scale,n_items", [
    (1, 752, 480, 8, 1.2, 400),
    (2, 752, 480, 4, 2.0, 300),        # EuRoC.yaml pyramid
    (3, 320, 240, 8, 1.2, 257),
])
def test_find_direct_projection_bit_exact(gpu, oracle, synth, seed, w, h, nlevels, scale, n_items):
    s, t, kfs, cur = _scene(gpu, synth, seed, w, h, nlevels, scale, n_items)
    pyr_bytes = sum(a.size + 16 for a in cur.refs)
    dm = gpu.DirectMatcher(s["K"], t["scale"], t["inv_scale"], t["inv_sigma2"][1], max_slots=len(kfs) + 1,
                           max_items=n_items, max_pyramid_bytes=pyr_bytes)
    for k, f in enumerate(kfs):
        dm.upload_frame(k, f)
    dm.upload_frame(len(kfs), cur)
    px, lvl, ok = dm.FindDirectProjection(len(kfs), s["items"])
    opx, olvl, ook = oracle.find_direct_projection(kfs, cur, s["K"], t["scale"], t["inv_scale"], t["inv_sigma2"][1], s["items"])
    assert (lvl == olvl).all()
    assert (ok == ook).all(), "return value differs on %d items" % (ok != ook).sum()
    bad = np.nonzero((px.view(np.uint32) != opx.view(np.uint32)).any(1))[0]
    assert bad.size == 0, "px differs on %d items, e.g. %s vs %s" % (bad.size, px[bad[:3]], opx[bad[:3]])
    assert ok.mean() > 0.4 and lvl.max() >= 1          # the scene exercises success, failure and search levels > 0


def test_pose_update_and_errors(gpu, oracle, synth):
    s, t, kfs, cur = _scene(gpu, synth, 5, 320, 240, 4, 1.5, 64)
    dm = gpu.DirectMatcher(s["K"], t["scale"], t["inv_scale"], t["inv_sigma2"][1], max_slots=4, max_items=64,
                           max_pyramid_bytes=2 * 320 * 240)
    for k, f in enumerate(kfs):
        dm.upload_frame(k, f)
    with pytest.raises(gpu.YgzError):                  # current-frame slot not uploaded yet
        dm.FindDirectProjection(3, s["items"])
    dm.upload_frame(3, cur)
    # a different current pose: set_pose on the resident slot == a fresh frame struct with that pose for the oracle
    q2 = synth.quat_from_rotvec([0.004, -0.003, 0.002]); t2 = np.array([0.01, -0.02, 0.015], np.float32)
    dm.set_pose(3, q2, t2)
    cur2 = gpu.make_direct_frame(cur.refs, q2, t2)
    px, lvl, ok = dm.FindDirectProjection(3, s["items"])
    opx, olvl, ook = oracle.find_direct_projection(kfs, cur2, s["K"], t["scale"], t["inv_scale"], t["inv_sigma2"][1], s["items"])
    assert (lvl == olvl).all() and (ok == ook).all() and (px.view(np.uint32) == opx.view(np.uint32)).all()
    bad = s["items"].copy(); bad["octave"][0] = 9
    with pytest.raises(gpu.YgzError):
        dm.FindDirectProjection(3, bad)
