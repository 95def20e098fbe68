"""The oracle's Frame::ComputeStereoMatches restatement (oracle/stereo_matches.cpp) against the REFERENCE's own function: the text of
src/Frame.cc:509-682 compiled, as it stands, as a member of the stand-in Frame (oracle/_ref/libstereo_ref.so: oracle/Makefile cuts the
function out of the reference source at build time, oracle/ref_stereo_shim.cpp).  mvuRight / mvDepth bit for bit, counts equal.
Build container only."""
import numpy as np
import pytest

import ref_lib


@pytest.fixture(scope="module")
def have_ref():
    if ref_lib._load("libstereo_ref.so") is None:
        pytest.skip("oracle/_ref/libstereo_ref.so not built (no /root/reference here)")


@pytest.mark.parametrize("seed,w,h,nf,mb,mbf", [(1, 752, 480, 1000, 0.54, 386.1448), (2, 1242, 375, 2000, 0.54, 386.1448), (3, 320, 240, 400, 0.11, 47.9),
                                                (4, 640, 360, 800, 0.5, 1.0)])
def test_compute_stereo_matches(have_ref, pkg, oracle, synth, seed, w, h, nf, mb, mbf):
    left, right, band, disp = synth.synth_stereo_pair(seed, w, h)
    oxL, oxR = oracle.Extractor(nf, 1.2, 8, 20, 7), oracle.Extractor(nf, 1.2, 8, 20, 7)
    kL, dL = oxL.extract(left)
    kR, dR = oxR.extract(right)
    t = oxL.tables()
    pyrL = pkg.make_direct_frame([oxL.level(l) for l in range(8)])
    pyrR = pkg.make_direct_frame([oxR.level(l) for l in range(8)])
    ou, oz, on = oracle.compute_stereo_matches(pyrL, pyrR, kL, dL, kR, dR, t["scale"], t["inv_scale"], mb, mbf)
    ru, rz, rn = ref_lib.compute_stereo_matches(pyrL, pyrR, kL, dL, kR, dR, t["scale"], t["inv_scale"], mb, mbf)
    assert on == rn
    assert (ou.view(np.uint32) == ru.view(np.uint32)).all(), int((ou != ru).sum())
    assert (oz.view(np.uint32) == rz.view(np.uint32)).all()
    if mbf > 40:
        assert on > 0.3 * len(kL)
    else:
        assert on == 0                     # maxD = mbf / mb = 2 px: every true disparity is out of range
