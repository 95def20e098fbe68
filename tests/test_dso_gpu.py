"""GPU parity of ORBextractor::operator()(Frame*, ..., DSO_KEYPOINT) (src/ORBextractor.cc:1031-1127 with
ComputeKeyPointsDSOSingleLevel :1277-1381, ShiTomasiScore :1152-1186, Thirdparty/fast FAST-10) against the oracle, whose
FAST-10 is pinned to the reference's own compiled code (tests/test_fast10_oracle.py): key points (all 7 fields, order
included), refreshed angles of the pre-existing keys, descriptors and the adaptive grid size carried across frames."""
import numpy as np
import pytest

from test_extractor_gpu import assert_kps_equal

pytestmark = pytest.mark.gpu


def _existing(oracle_ex, img, w, h, step):
    base, _ = oracle_ex.extract(img)
    ex = base[base["octave"] <= 1][::step].copy()
    keep = (ex["x"] > 30) & (ex["y"] > 30) & (ex["x"] < w - 30) & (ex["y"] < h - 30)
    return ex[keep]


@pytest.mark.parametrize("w,h,nf,seed", [(752, 480, 1000, 3), (320, 240, 300, 5), (640, 360, 150, 7)])
def test_dso_stream_bit_exact(gpu, oracle, synth, w, h, nf, seed):
    frames = synth.synth_stream(seed, 4, w, h)
    ex = gpu.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h)
    ox = oracle.Extractor(nf, 1.2, 8, 20, 7)
    gg, og = -1, -1
    grids = []
    for i in range(4):
        existing = _existing(ox, frames[i], w, h, 7) if i else np.zeros(0, gpu.KP_DTYPE)
        okps, odesc, og = ox.extract_frame_dso(frames[i], existing, og)
        kps, desc, gg = ex.extract_frame_dso(frames[i], existing, gg)
        assert gg == og, "mnGridSize after frame %d: %d vs %d" % (i, gg, og)
        assert_kps_equal(kps, okps, "frame %d" % i)
        assert (desc == odesc).all(), "descriptors of frame %d: %d rows differ" % (i, (desc != odesc).any(1).sum())
        assert len(kps) - len(existing) >= min(nf, 50)
        grids.append(gg)
    assert grids[-1] != int(np.sqrt(1.0 * h * w / nf))          # the adaptive cell size moved away from its initial value


def test_dso_cell_size_paths(gpu, oracle, synth):
    """Cell sizes below 22 px use the plain FAST-10 region (whole cell), others the SSE2 region (interior); both, the
    threshold-5 fallback (low-contrast image) and the no-corner frame are exercised by forcing mnGridSize."""
    w, h = 480, 360
    a = synth.synth_frame(13, w, h)
    low = (a.astype(np.int32) * 30 // 255 + 110).astype(np.uint8)
    flat = np.full((h, w), 77, np.uint8)
    ex = gpu.ORBextractor(400, 1.2, 8, 20, 7, width=w, height=h)
    ox = oracle.Extractor(400, 1.2, 8, 20, 7)
    none = np.zeros(0, gpu.KP_DTYPE)
    for img in (a, low):
        for g0 in (9, 19, 21, 22, 31, 47):
            okps, odesc, og = ox.extract_frame_dso(img, none, g0)
            kps, desc, gg = ex.extract_frame_dso(img, none, g0)
            assert gg == og and len(kps) == len(okps), (g0, gg, og, len(kps), len(okps))
            assert_kps_equal(kps, okps, "grid %d" % g0)
            assert (desc == odesc).all()
    kps, desc, gg = ex.extract_frame_dso(flat, none, -1)
    okps, odesc, og = ox.extract_frame_dso(flat, none, -1)
    assert len(kps) == 0 == len(okps) and gg == og
