"""CPU tests (no GPU): every OpenCV primitive restated in oracle/cvprims.cpp against cv2 itself, live, on fresh
seeds, plus the composed pyramid / per-cell FAST / blur pipeline (skipped where cv2 is not importable)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
cv2.setNumThreads(1)


def cvfast(img, t):
    f = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    return np.array([(p.pt[0], p.pt[1], p.response) for p in f.detect(np.ascontiguousarray(img))], np.float32).reshape(-1, 3)


@pytest.mark.parametrize("seed,w,h", [(101, 752, 480), (102, 401, 233), (103, 64, 80)])
def test_resize_blur_border_fast(oracle, synth, seed, w, h):
    img = synth.synth_frame(seed, max(w, 64), max(h, 64))[:h, :w]
    img = np.ascontiguousarray(img)
    for s in (1.2, 2.0, 1.37):
        dw, dh = int(round(w / s)), int(round(h / s))
        assert (oracle.resize(img, dw, dh) == cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)).all()
    assert (oracle.gauss7(img) == cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)).all()
    assert (oracle.border101(img, 19) == cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)).all()
    for t in (20, 7, 0, 60):
        k = oracle.fast9(img, t)
        ref = cvfast(img, t)
        assert len(k) == len(ref), t
        if len(k):
            assert (k["x"] == ref[:, 0]).all() and (k["y"] == ref[:, 1]).all() and (k["response"] == ref[:, 2]).all()


def test_fast_on_strided_roi(oracle, synth):
    img = synth.synth_frame(7, 320, 240)
    for (y0, y1, x0, x1) in [(16, 58, 16, 52), (100, 142, 200, 239), (0, 7, 0, 7), (10, 16, 10, 40)]:
        roi = img[y0:y1, x0:x1]                       # non-contiguous view: stride != width
        k = oracle.fast9(roi, 20)
        ref = cvfast(roi, 20)
        assert len(k) == len(ref)
        if len(k):
            assert (k["x"] == ref[:, 0]).all() and (k["y"] == ref[:, 1]).all() and (k["response"] == ref[:, 2]).all()


def test_fast_atan2(oracle):
    rng = np.random.default_rng(5)
    y = rng.integers(-2000000, 2000000, 100000).astype(np.float32)
    x = rng.integers(-2000000, 2000000, 100000).astype(np.float32)
    y[:6] = [0, 0, 3, -3, 1, 1e-30]
    x[:6] = [0, -2, 0, 0, 1, 1e-30]
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    assert (oracle.fast_atan2(y, x) == ref).all()


def test_pyramid_cells_and_blur_vs_cv2_composition(oracle, synth):
    """ComputePyramid + the per-cell cv::FAST(20)/cv::FAST(7) loop + per-level GaussianBlur, composed from cv2
    exactly as src/ORBextractor.cc:725-804,1009-1010,1129-1150 does, equals the oracle's stage outputs."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import cv2_candidates
    img = synth.synth_frame(77, 500, 300)
    ox = oracle.Extractor(700, 1.2, 8, 20, 7)
    ox.extract(img)
    lvl = img
    for l in range(8):
        lw, lh = ox.level_size(l)
        if l:
            lvl = cv2.resize(lvl, (lw, lh), interpolation=cv2.INTER_LINEAR)
        assert (lvl == ox.level(l)).all()
        assert (cv2.GaussianBlur(lvl, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101) == ox.level(l, blurred=True)).all()
        c, oc = cv2_candidates(lvl), ox.candidates(l)
        assert len(c) == len(oc)
        if len(c):
            assert (c[:, 0] == oc["x"]).all() and (c[:, 1] == oc["y"]).all() and (c[:, 2] == oc["response"]).all()
