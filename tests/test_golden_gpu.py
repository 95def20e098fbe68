"""GPU (through the C ABI) against the committed golden fixtures of tests/golden/ — independent of the live oracle."""
import os
import numpy as np
import pytest
import helpers as H

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("tag", ["c1_752x480_seed1", "c5_1242x375_seed5", "c3pyr_752x480_seed7", "small_320x240_seed3"])
def test_extractor_vs_fixture(gpu, synth, tag):
    g = np.load(os.path.join(GOLD, "extract_%s.npz" % tag))
    w, h = int(g["w"]), int(g["h"])
    img = synth.synth_frame(int(g["seed"]), w, h)
    ex = gpu.ORBextractor(int(g["nfeatures"]), float(g["scale"]), int(g["nlevels"]), 20, 7, width=w, height=h)
    kps, desc = ex(img)
    assert len(kps) == len(g["kps"])
    for f in kps.dtype.names:
        assert (kps[f] == g["kps"][f]).all(), f
    assert (desc == g["desc"]).all()
    assert [len(ex.candidates(0, l)) for l in range(int(g["nlevels"]))] == list(g["n_candidates"])


def test_matcher_vs_fixture(gpu):
    g = np.load(os.path.join(GOLD, "matcher_480x360.npz"))
    ka, da, kb, db, mp = g["ka"], g["da"], g["kb"], g["db"], g["mp"]
    m = gpu.ORBmatcher(0.7, True)
    kf = gpu.make_frame_view(ka, da, map_point=mp, node_of_feature=H.synth_vocab_nodes(da))
    f = gpu.make_frame_view(kb, db, node_of_feature=H.synth_vocab_nodes(db))
    mps = gpu.make_map_points(np.zeros((len(ka), 3)), da)
    n, out = m.SearchByBoW(kf, mps, f)
    assert n == int(g["n_bow"]) and (out == g["bow_out"]).all()
    bi, bd, sd = m.bf_top2(da, db)
    assert (bi == g["bf_idx"]).all() and (bd == g["bf_d"]).all() and (sd == g["bf_s"]).all()
    prev = np.stack([ka["x"], ka["y"]], -1).astype(np.float32)
    f1 = gpu.make_frame_view(ka, da, bounds=(0, 0, 480, 360))
    f2 = gpu.make_frame_view(kb, db, bounds=(0, 0, 480, 360))
    n, m12, prev_out = gpu.ORBmatcher(0.9, True).SearchForInitialization(f1, f2, prev, 100)
    assert n == int(g["n_init"]) and (m12 == g["init_m12"]).all() and (prev_out == g["init_prev"]).all()


def test_sia_vs_fixture(gpu, oracle, synth):
    import test_sia_gpu as T
    g = np.load(os.path.join(GOLD, "sia_752x480.npz"))
    sia = gpu.SparseImgAlign(3, 1, 10)
    for seed in (0, 1):
        pg, s = T.build_problem(gpu, oracle, synth, seed, gpu.SiaProblem)     # oracle only builds the input pyramid here
        n, q, t, _ = sia.run(pg)
        assert n == int(g["n%d" % seed])
        assert T.pose_diff(q, t, g["q%d" % seed], g["t%d" % seed]) <= 1e-4
