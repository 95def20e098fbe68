"""GPU parity of Frame::ComputeStereoMatches (src/Frame.cc:509-682) against the oracle: mvuRight and mvDepth bit-exact
(every float operation of the path is restated with explicit rounding), and the recovered disparities match the
synthetic pair's ground truth."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,w,h,nf", [(1, 752, 480, 1000), (2, 1242, 375, 2000), (3, 320, 240, 400)])
def test_compute_stereo_matches_bit_exact(gpu, oracle, synth, seed, w, h, nf):
    left, right, band, disp = synth.synth_stereo_pair(seed, w, h)
    exL = gpu.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h)
    exR = gpu.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h)
    (kL, dL), = exL.extract_batch(left[None])
    (kR, dR), = exR.extract_batch(right[None])
    t = exL.tables()
    mbf, mb = 386.1448, 0.54                      # KITTI-like bf and baseline: maxD = mbf / mb = 715 px
    u, z, n = gpu.ComputeStereoMatches(exL, exR, kL, dL, kR, dR, mb, mbf)
    pyrL = gpu.make_direct_frame([exL.level(0, l) for l in range(8)])
    pyrR = gpu.make_direct_frame([exR.level(0, l) for l in range(8)])
    ou, oz, on = oracle.compute_stereo_matches(pyrL, pyrR, kL, dL, kR, dR, t["scale"], t["inv_scale"], mb, mbf)
    assert n == on
    assert (u.view(np.uint32) == ou.view(np.uint32)).all(), "mvuRight differs on %d keys" % (u != ou).sum()
    assert (z.view(np.uint32) == oz.view(np.uint32)).all()
    m = u >= 0
    assert m.sum() > 0.3 * len(kL)
    want = np.array(disp, np.float32)[band[np.clip(np.rint(kL["y"][m]).astype(int), 0, h - 1)]]
    got = kL["x"][m] - u[m]
    assert np.median(np.abs(got - want)) < 0.6 and (np.abs(got - want) < 2.5).mean() > 0.85


def test_stereo_edge_cases(gpu, synth):
    w, h = 320, 240
    left, right, _, _ = synth.synth_stereo_pair(4, w, h)
    exL = gpu.ORBextractor(300, 1.2, 8, 20, 7, width=w, height=h)
    exR = gpu.ORBextractor(300, 1.2, 8, 20, 7, width=w, height=h)
    (kL, dL), = exL.extract_batch(left[None])
    (kR, dR), = exR.extract_batch(right[None])
    u, z, n = gpu.ComputeStereoMatches(exL, exR, kL, dL, kR[:0], dR[:0], 0.5, 40.0)       # no right keys
    assert n == 0 and (u == -1).all() and (z == -1).all()
    u, z, n = gpu.ComputeStereoMatches(exL, exR, kL, dL, kR, dR, 0.5, 1.0)                # maxD = 2 px: disparities of 6+ are out of range
    assert n == 0 and (u == -1).all()


@pytest.mark.parametrize("w,h,nf,B", [(1242, 375, 2000, 5), (640, 360, 700, 3)])
def test_stereo_batch_device_resident(gpu, oracle, synth, w, h, nf, B):
    """ygz_stereo_batch: B stereo pairs straight from the two extractors' device outputs (association without the host row table,
    median cut on the device, dense brute force) against the oracle run pair by pair — bit-exact, counts included."""
    pairs = [synth.synth_stereo_pair(10 + i, w, h) for i in range(B)]
    left = np.stack([p[0] for p in pairs]); right = np.stack([p[1] for p in pairs])
    if B > 2:
        right[2] = 128                              # a textureless right image: no right keys for that pair -> no matches
    exL = gpu.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h, max_batch=B)
    exR = gpu.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h, max_batch=B)
    resL = exL.extract_batch(left)
    resR = exR.extract_batch(right)
    t = exL.tables()
    mbf, mb = 386.1448, 0.54
    out = gpu.StereoBatch(exL, exR, B, mb, mbf, brute_force=True)
    for f in range(B):
        (kL, dL), (kR, dR) = resL[f], resR[f]
        n = len(kL)
        pyrL = gpu.make_direct_frame([exL.level(f, l) for l in range(8)])
        pyrR = gpu.make_direct_frame([exR.level(f, l) for l in range(8)])
        ou, oz, on = oracle.compute_stereo_matches(pyrL, pyrR, kL, dL, kR, dR, t["scale"], t["inv_scale"], mb, mbf)
        assert out["n_matches"][f] == on, (f, out["n_matches"][f], on)
        assert (out["u_right"][f, :n].view(np.uint32) == ou.view(np.uint32)).all()
        assert (out["depth"][f, :n].view(np.uint32) == oz.view(np.uint32)).all()
        assert (out["u_right"][f, n:] == -1).all() and (out["depth"][f, n:] == -1).all()
        if len(kR):
            bi, bd, sd = oracle.bf_top2(dL, dR)
            assert (out["bf_best"][f, :n] == bi).all() and (out["bf_dist"][f, :n] == bd).all() and (out["bf_second"][f, :n] == sd).all()
        assert (out["bf_best"][f, n:] == -1).all()
        if f == 2:
            assert on == 0 and len(kR) == 0
    assert out["n_matches"][0] > 0.3 * len(resL[0][0])
    # the per-frame entry point on the same handles gives the same answer
    u, z, n1 = gpu.ComputeStereoMatches(exL, exR, resL[1][0], resL[1][1], resR[1][0], resR[1][1], mb, mbf, frame=1)
    assert n1 == out["n_matches"][1] and (u.view(np.uint32) == out["u_right"][1, :len(u)].view(np.uint32)).all()
    ms, stages = gpu.time_stereo_frontend(exL, exR, B, 2, mb, mbf)
    assert ms > 0 and stages["match"] > 0
