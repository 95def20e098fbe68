"""Pins the oracle's FAST-10 restatement (oracle/dso_detect.cpp: fast10_detect_cell) against the REFERENCE's own
Thirdparty/fast code compiled from where it lies (oracle/_ref/libfast_ref.so built by oracle/Makefile): identical
corner lists (order included) for the plain detector (window width < 22: every pixel of the window) and the SSE2
detector (interior [3, w-3) x [3, h-3)), thresholds 5 / 20 / 75, window sizes the DSO key-point selection uses (7..40)
and whole images, plus the upstream known answer (167 corners on test1.png, threshold 75) where the PNG exists."""
import ctypes as C
import os
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libfast_ref.so")
PNG = "/root/reference/Thirdparty/fast/test/data/test1.png"


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/libfast_ref.so not built (needs /root/reference at build time)")
    return C.CDLL(SO)


def _ref_detect(ref, img, x0, y0, w, h, t):
    xy = np.zeros(2 * 40000, np.int16)
    sub = img[y0:, x0:]
    n = ref.fast_ref_detect10_sse2(C.c_void_p(sub.ctypes.data), w, h, img.shape[1], t, xy.ctypes.data_as(C.c_void_p), 40000)
    return xy[:2 * n].reshape(-1, 2).copy()


def _orc_detect(oracle, img, x0, y0, w, h, t):
    xy = np.zeros(2 * 40000, np.int16)
    sub = img[y0:, x0:]
    L = oracle.lib()
    L.orc_fast10_detect_cell.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    n = L.orc_fast10_detect_cell(C.c_void_p(sub.ctypes.data), w, h, img.shape[1], t, xy.ctypes.data_as(C.c_void_p), 40000)
    return xy[:2 * n].reshape(-1, 2).copy()


def test_cells_match_reference(ref, oracle, synth):
    rng = np.random.default_rng(3)
    imgs = [synth.synth_frame(31, 320, 240), rng.integers(0, 256, (240, 320), dtype=np.uint8),
            (rng.integers(0, 2, (240, 320)) * 255).astype(np.uint8)]
    total = 0
    for img in imgs:
        img = np.ascontiguousarray(img)
        for g in (7, 12, 19, 21, 22, 24, 29, 37, 40):
            for t in (5, 20, 75):
                for _ in range(6):
                    x0 = int(rng.integers(8, 320 - g - 8)); y0 = int(rng.integers(8, 240 - g - 8))
                    a = _ref_detect(ref, img, x0, y0, g, g, t)
                    b = _orc_detect(oracle, img, x0, y0, g, g, t)
                    assert a.shape == b.shape and (a == b).all(), (g, t, x0, y0, len(a), len(b))
                    total += len(a)
    assert total > 2000


def test_whole_image_matches_reference(ref, oracle, synth):
    for seed, w, h in ((5, 752, 480), (6, 333, 211)):
        img = synth.synth_frame(seed, w, h)
        for t in (20, 40):
            a = _ref_detect(ref, img, 0, 0, w, h, t)
            b = _orc_detect(oracle, img, 0, 0, w, h, t)
            assert len(a) > 100 and a.shape == b.shape and (a == b).all()


def test_upstream_known_answer_through_the_oracle(ref, oracle):
    cv2 = pytest.importorskip("cv2")
    if not os.path.exists(PNG):
        pytest.skip("reference tree not present on this box")
    img = np.ascontiguousarray(cv2.resize(cv2.imread(PNG, 0), (752, 480)))
    assert len(_orc_detect(oracle, img, 0, 0, 752, 480, 75)) == 167 == len(_ref_detect(ref, img, 0, 0, 752, 480, 75))
