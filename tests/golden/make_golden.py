"""Generates the committed golden fixtures (run in the build container: needs cv2 4.13 = the pinned behaviour of the
OpenCV primitives the reference calls, and the oracle).  Usage: python tests/golden/make_golden.py

  prims_cv2.npz        outputs of cv2 itself (resize / GaussianBlur / FAST / fastAtan2 / copyMakeBorder) on a seeded image
  extract_*.npz        oracle ORBextractor outputs; at generation time the pyramid, the per-cell FAST candidates and the
                       blurred levels are cross-checked against a cv2-composed pipeline (src/ORBextractor.cc:725-804,1129-1150)
  matcher_*.npz        oracle ORBmatcher outputs on oracle features of two overlapping crops
  sia_*.npz            oracle SparseImgAlign poses / traces
"""
import hashlib
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cv2  # noqa: E402
import __graft_entry__ as g  # noqa: E402
import oracle_lib as O  # noqa: E402
import helpers as H  # noqa: E402

pkg = g.load_package()
synth = importlib.import_module("orb_ygz_slam_b200.synth")
cv2.setNumThreads(1)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cv2_fast(img, t):
    f = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    k = f.detect(np.ascontiguousarray(img))
    return np.array([(p.pt[0], p.pt[1], p.response) for p in k], np.float32).reshape(-1, 3)


def cv2_candidates(level_img, ini_th=20, min_th=7):
    """ComputeKeyPointsOctTree's cell loop (src/ORBextractor.cc:730-781) with cv2.FAST per cell."""
    h, w = level_img.shape
    minB, maxBX, maxBY = 16, w - 16, h - 16
    width, height = np.float32(maxBX - minB), np.float32(maxBY - minB)
    nCols, nRows = int(width / np.float32(30)), int(height / np.float32(30))
    out = []
    if nCols <= 0 or nRows <= 0:
        return np.zeros((0, 3), np.float32)
    wCell, hCell = int(np.ceil(width / nCols)), int(np.ceil(height / nRows))
    for i in range(nRows):
        iniY = minB + i * hCell
        maxY = iniY + hCell + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(nCols):
            iniX = minB + j * wCell
            maxX = iniX + wCell + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            cell = level_img[iniY:maxY, iniX:maxX]
            k = cv2_fast(cell, ini_th)
            if len(k) == 0:
                k = cv2_fast(cell, min_th)
            for x, y, r in k:
                out.append((x + j * wCell, y + i * hCell, r))
    return np.array(out, np.float32).reshape(-1, 3)


def make_prims():
    img = synth.synth_frame(42, 160, 120)
    rng = np.random.default_rng(0)
    ys = rng.integers(-300000, 300000, 4000).astype(np.float32)
    xs = rng.integers(-300000, 300000, 4000).astype(np.float32)
    ys[:8] = [0, 0, 1, -1, 5, -5, 0, 7]
    xs[:8] = [0, 1, 0, 0, 5, 5, -3, -7]
    at = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(ys, xs)], np.float32)
    np.savez_compressed(os.path.join(HERE, "prims_cv2.npz"), img=img,
                        resize_133x100=cv2.resize(img, (133, 100), interpolation=cv2.INTER_LINEAR),
                        resize_80x60=cv2.resize(img, (80, 60), interpolation=cv2.INTER_LINEAR),
                        resize_97x41=cv2.resize(img, (97, 41), interpolation=cv2.INTER_LINEAR),
                        blur=cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101),
                        border19=cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101),
                        fast20=cv2_fast(img, 20), fast7=cv2_fast(img, 7), atan_y=ys, atan_x=xs, atan=at,
                        cv2_version=np.array(cv2.__version__))


def make_extract(tag, seed, w, h, nf, sf, nl):
    img = synth.synth_frame(seed, w, h)
    ox = O.Extractor(nf, sf, nl, 20, 7)
    kps, desc = ox.extract(img)
    # cross-check against the cv2-composed pipeline
    lvl = img
    n_cand = []
    for l in range(nl):
        lw, lh = ox.level_size(l)
        if l > 0:
            lvl = cv2.resize(lvl, (lw, lh), interpolation=cv2.INTER_LINEAR)
        assert (lvl == ox.level(l)).all(), ("pyramid", l)
        assert (cv2.GaussianBlur(lvl, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101) == ox.level(l, blurred=True)).all(), ("blur", l)
        c = cv2_candidates(lvl)
        oc = ox.candidates(l)
        assert len(c) == len(oc) and (c[:, 0] == oc["x"]).all() and (c[:, 1] == oc["y"]).all() and (c[:, 2] == oc["response"]).all(), ("cand", l)
        n_cand.append(len(c))
    np.savez_compressed(os.path.join(HERE, "extract_%s.npz" % tag), seed=seed, w=w, h=h, nfeatures=nf, scale=sf, nlevels=nl,
                        img_sha=np.array(sha(img)), kps=kps, desc=desc, n_candidates=np.array(n_cand),
                        level_sha=np.array([sha(ox.level(l)) for l in range(nl)]))
    print(tag, len(kps), n_cand)


def make_matcher():
    (_, ka, da), (_, kb, db) = H.two_frames(O, synth, seed=1, w=480, h=360, nfeatures=600)
    mp = np.arange(len(ka), dtype=np.int32)
    mp[::5] = -1
    kf = pkg.make_frame_view(ka, da, map_point=mp, node_of_feature=H.synth_vocab_nodes(da), cls=O.FrameView)
    f = pkg.make_frame_view(kb, db, node_of_feature=H.synth_vocab_nodes(db), cls=O.FrameView)
    mps = pkg.make_map_points(np.zeros((len(ka), 3)), da, cls=O.MapPoints)
    out = np.zeros(len(kb), np.int32)
    n_bow = O.lib().orc_search_by_bow(0.7, 1, kf.ptr(), mps.ptr(), f.ptr(), out.ctypes.data_as(O.i32p))
    prev = np.stack([ka["x"], ka["y"]], -1).astype(np.float32)
    m12 = np.zeros(len(ka), np.int32)
    f1 = pkg.make_frame_view(ka, da, bounds=(0, 0, 480, 360), cls=O.FrameView)
    f2 = pkg.make_frame_view(kb, db, bounds=(0, 0, 480, 360), cls=O.FrameView)
    n_init = O.lib().orc_search_for_initialization(0.9, 1, f1.ptr(), f2.ptr(), prev.ctypes.data_as(O.f32p), m12.ctypes.data_as(O.i32p), 100)
    bi, bd, sd = O.bf_top2(da, db)
    np.savez_compressed(os.path.join(HERE, "matcher_480x360.npz"), ka=ka, da=da, kb=kb, db=db, mp=mp, bow_out=out, n_bow=n_bow,
                        init_m12=m12, n_init=n_init, init_prev=prev, bf_idx=bi, bf_d=bd, bf_s=sd)
    print("matcher", n_bow, n_init)


def make_sia():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_sia_gpu as T
    rec = {}
    for seed in (0, 1):
        po, s = T.build_problem(pkg, O, synth, seed, O.SiaProblem)
        n, q, t, Hm, tr = O.sia_run(po, 3, 1, 10, 1e-6)
        rec["q%d" % seed], rec["t%d" % seed], rec["n%d" % seed], rec["trace%d" % seed] = q, t, n, tr
        rec["ref_sha%d" % seed] = np.array(sha(s["ref"]))
        rec["cur_sha%d" % seed] = np.array(sha(s["cur"]))
        print("sia", seed, n, q, t)
    np.savez_compressed(os.path.join(HERE, "sia_752x480.npz"), **rec)


def make_direct():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_oracle_direct as T
    rec = {}
    for seed in (1, 2):
        s, t, kfs, cur = T.build_scene(pkg, O, synth, seed)
        px, lvl, ok = O.find_direct_projection(kfs, cur, s["K"], t["scale"], t["inv_scale"], t["inv_sigma2"][1], s["items"])
        rec["px%d" % seed], rec["lvl%d" % seed], rec["ok%d" % seed] = px, lvl, ok
        rec["cur_sha%d" % seed] = np.array(sha(s["cur_img"]))
        print("direct", seed, ok.mean(), np.bincount(lvl))
    np.savez_compressed(os.path.join(HERE, "direct_v1.npz"), **rec)


def make_dso_stereo_vocab():
    rec = {}
    # DSO key-point selection over a 3-frame stream (adaptive grid carried across frames)
    w, h, nf = 480, 360, 400
    frames = synth.synth_stream(11, 3, w, h)
    ox = O.Extractor(nf, 1.2, 8, 20, 7)
    base, _ = ox.extract(frames[0])
    existing = base[(base["octave"] == 0) & (base["x"] > 30) & (base["y"] > 30) & (base["x"] < w - 30) & (base["y"] < h - 30)][::5].copy()
    g = -1
    for i, f in enumerate(frames):
        kps, desc, g = ox.extract_frame_dso(f, existing, g)
        rec["dso_kps%d" % i], rec["dso_desc%d" % i], rec["dso_grid%d" % i] = kps, desc, g
    rec["dso_frame_sha"] = np.array(sha(frames))
    # stereo association on a synthetic rectified pair
    left, right, band, disp = synth.synth_stereo_pair(7, w, h, disparities=(5, 11, 19))
    oxL = O.Extractor(600, 1.2, 8, 20, 7); oxR = O.Extractor(600, 1.2, 8, 20, 7)
    kL, dL = oxL.extract(left); kR, dR = oxR.extract(right)
    t = oxL.tables()
    pyrL = pkg.make_direct_frame([oxL.level(l).copy() for l in range(8)])
    pyrR = pkg.make_direct_frame([oxR.level(l).copy() for l in range(8)])
    u, z, n = O.compute_stereo_matches(pyrL, pyrR, kL, dL, kR, dR, t["scale"], t["inv_scale"], 0.5, 100.0)
    rec["st_u"], rec["st_z"], rec["st_n"], rec["st_right_sha"] = u, z, n, np.array(sha(right))
    # vocabulary transform on a synthetic tree
    data = synth.synth_vocabulary(3, k=6, L=3)
    v = O.Vocabulary(data=data)
    q, tt, _ = synth.synth_descriptors(5, 400, 400)
    wid, wgt, nid = v.transform_features(np.concatenate([q, tt]), 2)
    ids, vals, fv = v.transform(q, 2)
    rec["voc_sha"] = np.array(sha(np.frombuffer(data, np.uint8)))
    rec["voc_wid"], rec["voc_w"], rec["voc_nid"], rec["voc_bow_ids"], rec["voc_bow_vals"] = wid, wgt, nid, ids, vals
    np.savez_compressed(os.path.join(HERE, "f_rows_v1.npz"), **rec)
    print("f rows", [len(rec["dso_kps%d" % i]) for i in range(3)], n, len(ids))


if __name__ == "__main__":
    make_prims()
    make_extract("c1_752x480_seed1", 1, 752, 480, 1000, 1.2, 8)
    make_extract("c5_1242x375_seed5", 5, 1242, 375, 2000, 1.2, 8)
    make_extract("c3pyr_752x480_seed7", 7, 752, 480, 1000, 2.0, 4)
    make_extract("small_320x240_seed3", 3, 320, 240, 500, 1.2, 8)
    make_matcher()
    make_sia()
    make_direct()
    make_dso_stereo_vocab()
