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


@pytest.mark.parametrize("w,h,nf", [(752, 480, 1000), (480, 360, 300)])
def test_dso_batch_equals_per_frame(gpu, oracle, synth, w, h, nf):
    """ygz_orb_extract_frame_dso_batch: B independent streams (own existing keys, own mnGridSize: first frame, forced small / large cells,
    a low-contrast stream that needs the threshold-5 fallback and several passes, a flat image that never finds a corner), two rounds with
    the grid sizes carried over — every frame equal to the oracle's per-frame result, and to the per-frame device entry."""
    B = 6
    imgs0 = [synth.synth_frame(40 + i, w, h) for i in range(B)]
    imgs0[3] = (imgs0[3].astype(np.int32) * 30 // 255 + 110).astype(np.uint8)
    imgs0[4] = np.full((h, w), 77, np.uint8)
    imgs1 = [np.ascontiguousarray(np.roll(im, 3, axis=1)) for im in imgs0]
    ex = gpu.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h, max_batch=B)
    one = gpu.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h)
    ox = oracle.Extractor(nf, 1.2, 8, 20, 7)
    g = np.array([-1, 9, 47, -1, -1, 22], np.int32)
    og = g.copy()
    none = np.zeros(0, gpu.KP_DTYPE)
    for rnd, imgs in enumerate((imgs0, imgs1)):
        existing = [(_existing(ox, imgs[f], w, h, 5 + f) if (rnd or f % 2) and f != 4 else none) for f in range(B)]
        res, g = ex.extract_frame_dso_batch(np.stack(imgs), existing, g)
        for f in range(B):
            okps, odesc, og[f] = ox.extract_frame_dso(imgs[f], existing[f], int(og[f]))
            kps, desc = res[f]
            assert g[f] == og[f], "round %d frame %d: mnGridSize %d vs %d" % (rnd, f, g[f], og[f])
            assert_kps_equal(kps, okps, "round %d frame %d" % (rnd, f))
            assert (desc == odesc).all(), "round %d frame %d descriptors" % (rnd, f)
        skps, sdesc, sg = one.extract_frame_dso(imgs[0], existing[0], -1 if rnd == 0 else int(prev_g0))
        assert sg == g[0] and (sdesc == res[0][1]).all()
        prev_g0 = g[0]
    assert len(res[4][0]) == 0 and len(res[3][0]) > 20
    # device-resident outputs agree with what was returned (zero-copy chaining)
    _, desc_d, cnt_d = ex.download(B)
    for f in range(B):
        assert cnt_d[f] == len(res[f][0]) and (desc_d[f, :cnt_d[f]] == res[f][1]).all()
    # argument errors
    with pytest.raises(Exception):
        ex.extract_frame_dso_batch(np.stack(imgs0), [none] * B, np.array([5] * B, np.int32))          # cell size below 7
    with pytest.raises(Exception):
        ex.extract_frame_dso_batch(np.stack(imgs0 + imgs0), [none] * (2 * B), np.full(2 * B, -1, np.int32))   # more frames than the handle holds
