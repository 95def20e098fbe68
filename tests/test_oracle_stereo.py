"""CPU check of the oracle's Frame::ComputeStereoMatches restatement (oracle/stereo_matches.cpp): recovers the known
per-band disparities of a synthetic rectified pair with sub-pixel accuracy (parity unpinned upstream)."""
import numpy as np


def test_oracle_recovers_disparity(oracle, pkg, synth):
    w, h = 480, 360
    left, right, band, disp = synth.synth_stereo_pair(7, w, h, disparities=(5, 11, 19))
    oxL = oracle.Extractor(600, 1.2, 8, 20, 7); oxR = oracle.Extractor(600, 1.2, 8, 20, 7)
    kL, dL = oxL.extract(left); kR, dR = oxR.extract(right)
    t = oxL.tables()
    pyrL = pkg.make_direct_frame([oxL.level(l).copy() for l in range(8)])
    pyrR = pkg.make_direct_frame([oxR.level(l).copy() for l in range(8)])
    u, z, n = oracle.compute_stereo_matches(pyrL, pyrR, kL, dL, kR, dR, t["scale"], t["inv_scale"], 0.5, 100.0)
    m = u >= 0
    assert n == m.sum() and n > 0.3 * len(kL)
    want = np.array(disp, np.float32)[band[np.clip(np.rint(kL["y"][m]).astype(int), 0, h - 1)]]
    got = kL["x"][m] - u[m]
    assert np.median(np.abs(got - want)) < 0.6 and (np.abs(got - want) < 2.5).mean() > 0.85
    assert np.allclose(z[m], 100.0 / np.maximum(got, 0.01), rtol=1e-5)
