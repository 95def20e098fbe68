"""CPU checks of the oracle's DSO key-point selection (oracle/dso_detect.cpp): Shi-Tomasi score against a float64
restatement, and the structural properties ComputeKeyPointsDSOSingleLevel guarantees (src/ORBextractor.cc:1277-1381)."""
import numpy as np


def test_shi_tomasi_matches_float64(oracle, synth):
    img = synth.synth_frame(2, 320, 240)
    ox = oracle.Extractor(300, 1.2, 8, 20, 7)
    ox.extract(img)
    L = ox.L
    rng = np.random.default_rng(0)
    I = img.astype(np.float64)
    for _ in range(200):
        u, v = int(rng.integers(6, 314)), int(rng.integers(6, 234))
        ys, xs = np.mgrid[v - 4:v + 4, u - 4:u + 4]
        dx = I[ys, xs + 1] - I[ys, xs - 1]
        dy = I[ys + 1, xs] - I[ys - 1, xs]
        a, b, c = (dx * dx).sum() / 128, (dy * dy).sum() / 128, (dx * dy).sum() / 128
        want = 0.5 * (a + b - np.sqrt((a + b) ** 2 - 4 * (a * b - c * c)))
        got = L.orc_shi_tomasi(ox.h, u, v)
        assert abs(got - want) <= 1e-3 * max(1.0, abs(want)), (u, v, got, want)
    assert L.orc_shi_tomasi(ox.h, 3, 100) == 0.0          # patch too close to the boundary (:1165-1166)


def test_dso_structure(oracle, synth):
    w, h, nf = 480, 360, 400
    frames = synth.synth_stream(11, 3, w, h)
    ox = oracle.Extractor(nf, 1.2, 8, 20, 7)
    base, _ = ox.extract(frames[0])
    existing = base[(base["octave"] == 0) & (base["x"] > 30) & (base["y"] > 30) & (base["x"] < w - 30) & (base["y"] < h - 30)][::5].copy()
    existing["angle"] = 0.0
    g = -1
    for f in frames:
        g_before = g if g > 0 else int(np.sqrt(1.0 * h * w / nf))
        kps, desc, g = ox.extract_frame_dso(f, existing, g)
        new = kps[len(existing):]
        assert len(new) >= nf and (new["octave"] == 0).all() and (new["size"] == 7).all() and (new["response"] == 0).all()
        assert new["x"].min() >= 20 and new["y"].min() >= 20 and new["x"].max() < w - 20 and new["y"].max() < h - 20
        occ = set(zip(np.rint(existing["x"]).astype(int), np.rint(existing["y"]).astype(int)))
        assert not occ & set(zip(new["x"].astype(int), new["y"].astype(int)))
        # at most three per cell of the final grid (g was raised by 5 afterwards because cnt > n)
        gf = g - 5
        assert g_before - gf in (0, 5, 10, 15) or gf == 7
        cells = (new["y"].astype(int) // gf) * 1000 + new["x"].astype(int) // gf
        assert np.bincount(np.unique(cells, return_inverse=True)[1]).max() <= 3
        assert (kps["angle"][:len(existing)] != 0).mean() > 0.9                # angles of the tracked keys are refreshed
