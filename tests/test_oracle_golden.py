"""CPU tests (no GPU): the oracle against the committed golden fixtures.
prims_cv2.npz holds outputs of cv2 4.13 itself (the pinned behaviour of the OpenCV primitives the reference calls);
extract_/matcher_/sia_ fixtures hold oracle outputs that were cross-checked against a cv2-composed pipeline when
they were generated (tests/golden/make_golden.py)."""
import hashlib
import os
import numpy as np
import pytest
import helpers as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_primitives_match_cv2_fixture(oracle):
    g = np.load(os.path.join(GOLD, "prims_cv2.npz"))
    img = g["img"]
    for key, (dw, dh) in (("resize_133x100", (133, 100)), ("resize_80x60", (80, 60)), ("resize_97x41", (97, 41))):
        assert (oracle.resize(img, dw, dh) == g[key]).all(), key
    assert (oracle.gauss7(img) == g["blur"]).all()
    assert (oracle.border101(img, 19) == g["border19"]).all()
    for t, key in ((20, "fast20"), (7, "fast7")):
        k = oracle.fast9(img, t)
        ref = g[key]
        assert len(k) == len(ref)
        assert (k["x"] == ref[:, 0]).all() and (k["y"] == ref[:, 1]).all() and (k["response"] == ref[:, 2]).all()
    assert (oracle.fast_atan2(g["atan_y"], g["atan_x"]) == g["atan"]).all()


@pytest.mark.parametrize("tag", ["c1_752x480_seed1", "c5_1242x375_seed5", "c3pyr_752x480_seed7", "small_320x240_seed3"])
def test_extractor_fixture(oracle, synth, tag):
    g = np.load(os.path.join(GOLD, "extract_%s.npz" % tag))
    img = synth.synth_frame(int(g["seed"]), int(g["w"]), int(g["h"]))
    assert sha(img) == str(g["img_sha"]), "synthetic generator drifted"
    ox = oracle.Extractor(int(g["nfeatures"]), float(g["scale"]), int(g["nlevels"]), 20, 7)
    kps, desc = ox.extract(img)
    assert len(kps) == len(g["kps"])
    for f in kps.dtype.names:
        assert (kps[f] == g["kps"][f]).all(), f
    assert (desc == g["desc"]).all()
    assert [len(ox.candidates(l)) for l in range(int(g["nlevels"]))] == list(g["n_candidates"])
    assert [sha(ox.level(l)) for l in range(int(g["nlevels"]))] == list(g["level_sha"])


def test_matcher_fixture(oracle, pkg):
    g = np.load(os.path.join(GOLD, "matcher_480x360.npz"))
    ka, da, kb, db, mp = g["ka"], g["da"], g["kb"], g["db"], g["mp"]
    kf = pkg.make_frame_view(ka, da, map_point=mp, node_of_feature=H.synth_vocab_nodes(da), cls=oracle.FrameView)
    f = pkg.make_frame_view(kb, db, node_of_feature=H.synth_vocab_nodes(db), cls=oracle.FrameView)
    mps = pkg.make_map_points(np.zeros((len(ka), 3)), da, cls=oracle.MapPoints)
    out = np.zeros(len(kb), np.int32)
    n = oracle.lib().orc_search_by_bow(0.7, 1, kf.ptr(), mps.ptr(), f.ptr(), out.ctypes.data_as(oracle.i32p))
    assert n == int(g["n_bow"]) and (out == g["bow_out"]).all()
    prev = np.stack([ka["x"], ka["y"]], -1).astype(np.float32)
    m12 = np.zeros(len(ka), np.int32)
    f1 = pkg.make_frame_view(ka, da, bounds=(0, 0, 480, 360), cls=oracle.FrameView)
    f2 = pkg.make_frame_view(kb, db, bounds=(0, 0, 480, 360), cls=oracle.FrameView)
    n = oracle.lib().orc_search_for_initialization(0.9, 1, f1.ptr(), f2.ptr(), prev.ctypes.data_as(oracle.f32p),
                                                   m12.ctypes.data_as(oracle.i32p), 100)
    assert n == int(g["n_init"]) and (m12 == g["init_m12"]).all() and (prev == g["init_prev"]).all()
    bi, bd, sd = oracle.bf_top2(da, db)
    assert (bi == g["bf_idx"]).all() and (bd == g["bf_d"]).all() and (sd == g["bf_s"]).all()
    # brute-force distances == numpy popcount
    d = H.hamming_matrix(da[:50], db[:60])
    assert (d.min(1) == bd[:50]).all() or True       # bf over all of db: only check the restricted matrix below
    bi2, bd2, sd2 = oracle.bf_top2(da[:50], db[:60])
    assert (d.min(1) == bd2).all() and (d.argmin(1) == bi2).all()


def test_sia_fixture(oracle, pkg, synth):
    import test_sia_gpu as T
    g = np.load(os.path.join(GOLD, "sia_752x480.npz"))
    for seed in (0, 1):
        po, s = T.build_problem(pkg, oracle, synth, seed, oracle.SiaProblem)
        assert sha(s["ref"]) == str(g["ref_sha%d" % seed]) and sha(s["cur"]) == str(g["cur_sha%d" % seed])
        n, q, t, Hm, tr = oracle.sia_run(po, 3, 1, 10, 1e-6)
        assert n == int(g["n%d" % seed])
        # same compiler flags, same arithmetic: the oracle reproduces its own fixture to float round-off
        assert np.abs(q - g["q%d" % seed]).max() < 1e-6 and np.abs(t - g["t%d" % seed]).max() < 1e-6
        assert len(tr) == len(g["trace%d" % seed])
        assert T.pose_diff(q, t, s["q_true"], s["t_true"]) < 5e-3


def test_f_rows_fixture(oracle, pkg, synth):
    """Oracle regression pins of the SURVEY 8(f) rows without their own fixture file: DSO key points over a stream,
    ComputeStereoMatches, vocabulary transform (tests/golden/f_rows_v1.npz, made by make_golden.py)."""
    g = np.load(os.path.join(GOLD, "f_rows_v1.npz"))
    w, h, nf = 480, 360, 400
    frames = synth.synth_stream(11, 3, w, h)
    assert sha(frames) == str(g["dso_frame_sha"])
    ox = oracle.Extractor(nf, 1.2, 8, 20, 7)
    base, _ = ox.extract(frames[0])
    existing = base[(base["octave"] == 0) & (base["x"] > 30) & (base["y"] > 30) & (base["x"] < w - 30) & (base["y"] < h - 30)][::5].copy()
    grid = -1
    for i, f in enumerate(frames):
        kps, desc, grid = ox.extract_frame_dso(f, existing, grid)
        assert grid == int(g["dso_grid%d" % i]) and len(kps) == len(g["dso_kps%d" % i])
        for fld in ("x", "y", "size", "angle", "response", "octave", "class_id"):
            assert (kps[fld] == g["dso_kps%d" % i][fld]).all(), (i, fld)
        assert (desc == g["dso_desc%d" % i]).all()
    left, right, band, disp = synth.synth_stereo_pair(7, w, h, disparities=(5, 11, 19))
    assert sha(right) == str(g["st_right_sha"])
    oxL = oracle.Extractor(600, 1.2, 8, 20, 7); oxR = oracle.Extractor(600, 1.2, 8, 20, 7)
    kL, dL = oxL.extract(left); kR, dR = oxR.extract(right)
    t = oxL.tables()
    pyrL = pkg.make_direct_frame([oxL.level(l).copy() for l in range(8)])
    pyrR = pkg.make_direct_frame([oxR.level(l).copy() for l in range(8)])
    u, z, n = oracle.compute_stereo_matches(pyrL, pyrR, kL, dL, kR, dR, t["scale"], t["inv_scale"], 0.5, 100.0)
    assert n == int(g["st_n"]) and (u.view(np.uint32) == g["st_u"].view(np.uint32)).all() and (z.view(np.uint32) == g["st_z"].view(np.uint32)).all()
    data = synth.synth_vocabulary(3, k=6, L=3)
    assert sha(np.frombuffer(data, np.uint8)) == str(g["voc_sha"])
    v = oracle.Vocabulary(data=data)
    q, tt, _ = synth.synth_descriptors(5, 400, 400)
    wid, wgt, nid = v.transform_features(np.concatenate([q, tt]), 2)
    ids, vals, fv = v.transform(q, 2)
    assert (wid == g["voc_wid"]).all() and (nid == g["voc_nid"]).all() and (wgt.view(np.uint32) == g["voc_w"].view(np.uint32)).all()
    assert (ids == g["voc_bow_ids"]).all() and (vals == g["voc_bow_vals"]).all()
