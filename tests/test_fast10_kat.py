"""The single upstream known-answer (Thirdparty/fast/test/test.cpp:6-10,40-54): the reference's own SSE2 FAST-10,
compiled from /root/reference into oracle/_ref/libfast_ref.so, finds 167 corners at threshold 75 on test1.png resized
to 752x480.  Runs only where the reference tree (image + sources) is present, i.e. in the build container."""
import ctypes as C
import os
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PNG = "/root/reference/Thirdparty/fast/test/data/test1.png"
SO = os.path.join(ROOT, "oracle", "_ref", "libfast_ref.so")


def test_reference_fast10_sse2_known_answer():
    cv2 = pytest.importorskip("cv2")
    if not os.path.exists(PNG) or not os.path.exists(SO):
        pytest.skip("reference tree / oracle/_ref not present on this box")
    img = cv2.resize(cv2.imread(PNG, 0), (752, 480))
    img = np.ascontiguousarray(img)
    xy = np.zeros(20000, np.int16)
    n = C.CDLL(SO).fast_ref_detect10_sse2(img.ctypes.data_as(C.c_void_p), 752, 480, 752, 75, xy.ctypes.data_as(C.c_void_p), 10000)
    assert n == 167
