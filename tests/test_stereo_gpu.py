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
