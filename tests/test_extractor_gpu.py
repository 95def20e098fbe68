"""GPU parity of the ORB extractor (through the C ABI) against the oracle: bit-exact pyramid levels,
FAST candidates (order + response), keypoints (all 7 fields) and 32-byte descriptors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KP_FIELDS = ("x", "y", "size", "angle", "response", "octave", "class_id")


def assert_kps_equal(a, b, what=""):
    assert len(a) == len(b), "%s: count %d vs %d" % (what, len(a), len(b))
    for f in KP_FIELDS:
        bad = np.nonzero(a[f] != b[f])[0]
        assert bad.size == 0, "%s: field %s differs at %s: %s vs %s" % (what, f, bad[:5], a[f][bad[:5]], b[f][bad[:5]])


CONFIGS = [
    # (w, h, nfeatures, scale, nlevels, seeds)
    (752, 480, 1000, 1.2, 8, [1, 2]),       # C1
    (1242, 375, 2000, 1.2, 8, [5]),          # C5 (KITTI shape)
    (752, 480, 1000, 2.0, 4, [7]),           # C3 pyramid (EuRoC.yaml): level 3 has no FAST cells
    (320, 240, 500, 1.2, 8, [3, 4]),
    (200, 150, 300, 1.5, 5, [9]),
    (1920, 1080, 2000, 1.2, 8, [11]),       # full HD: 4x the C1 pixels, 5.8 k cells
    (641, 479, 700, 1.2, 8, [12]),           # odd sizes: row pitch padding on every level
]


@pytest.mark.parametrize("w,h,nf,sf,nl,seeds", CONFIGS)
def test_extract_stages_bit_exact(gpu, oracle, synth, w, h, nf, sf, nl, seeds):
    imgs = np.stack([synth.synth_frame(s, w, h) for s in seeds])
    ex = gpu.ORBextractor(nf, sf, nl, 20, 7, width=w, height=h, max_batch=len(seeds))
    res = ex.extract_batch(imgs)
    ox = oracle.Extractor(nf, sf, nl, 20, 7)
    t = ex.tables()
    ot = ox.tables()
    for k in ("scale", "inv_scale", "sigma2", "inv_sigma2", "features_per_level"):
        assert (t[k] == ot[k]).all(), k
    for f, seed in enumerate(seeds):
        okps, odesc = ox.extract(imgs[f])
        for l in range(nl):
            assert ex.level_size(l) == ox.level_size(l)
            assert (ex.level(f, l) == ox.level(l)).all(), "pyramid level %d" % l
            gc = ex.candidates(f, l)
            oc = ox.candidates(l)
            assert_kps_equal(gc, oc, "candidates level %d" % l)
            if len(ox.level_keys(l)):          # the reference blurs a level only if it holds key points (src/ORBextractor.cc:1003-1010)
                assert (ex.blurred_level(f, l) == ox.level(l, blurred=True)).all(), "blurred level %d" % l
        kps, desc = res[f]
        assert_kps_equal(kps, okps, "keypoints seed %d" % seed)
        assert (desc == odesc).all(), "descriptors: %d rows differ" % (desc != odesc).any(1).sum()


def test_flat_and_noise_images(gpu, oracle):
    w, h = 320, 240
    rng = np.random.default_rng(0)
    imgs = np.stack([np.full((h, w), 128, np.uint8),                        # no corners at all
                     rng.integers(0, 256, (h, w), dtype=np.uint8),          # corners everywhere
                     np.zeros((h, w), np.uint8)])
    imgs[2, ::2, ::2] = 255                                                 # checker dots: many equal scores
    ex = gpu.ORBextractor(500, 1.2, 8, 20, 7, width=w, height=h, max_batch=3)
    res = ex.extract_batch(imgs)
    ox = oracle.Extractor(500, 1.2, 8, 20, 7)
    for f in range(3):
        okps, odesc = ox.extract(imgs[f])
        assert_kps_equal(res[f][0], okps, "image %d" % f)
        assert (res[f][1] == odesc).all()
    assert len(res[0][0]) == 0


@pytest.mark.parametrize("ini,mn", [(20, 7), (60, 7), (120, 3), (7, 20), (12, 12)])
def test_threshold_fallback(gpu, oracle, synth, ini, mn):
    """cv::FAST(iniTh) then, for empty cells only, cv::FAST(minTh) (src/ORBextractor.cc:765-770): high iniTh makes most
    cells take the fallback, iniTh < minTh makes the fallback a no-op; low-contrast image exercises both."""
    w, h = 320, 240
    a = synth.synth_frame(17, w, h)
    b = (a.astype(np.int32) * 40 // 255 + 100).astype(np.uint8)            # low contrast: few corners at iniTh
    imgs = np.stack([a, b])
    ex = gpu.ORBextractor(500, 1.2, 8, ini, mn, width=w, height=h, max_batch=2)
    res = ex.extract_batch(imgs)
    ox = oracle.Extractor(500, 1.2, 8, ini, mn)
    for f in range(2):
        okps, odesc = ox.extract(imgs[f])
        for l in range(8):
            assert_kps_equal(ex.candidates(f, l), ox.candidates(l), "candidates level %d" % l)
        assert_kps_equal(res[f][0], okps, "image %d" % f)
        assert (res[f][1] == odesc).all()


def test_border_pyramid(gpu, oracle, synth):
    w, h = 320, 240
    img = synth.synth_frame(11, w, h)
    ex = gpu.ORBextractor(500, 1.2, 8, 20, 7, width=w, height=h)
    ex.ComputePyramid(img)
    ox = oracle.Extractor(500, 1.2, 8, 20, 7)
    ox.extract(img)
    for l in (0, 3, 7):
        assert (ex.level(0, l, border=19) == oracle.border101(ox.level(l), 19)).all()


def test_extract_frame_existing_keys(gpu, oracle, synth):
    w, h = 752, 480
    img = synth.synth_frame(21, w, h)
    ex = gpu.ORBextractor(1000, 1.2, 8, 20, 7, width=w, height=h)
    ox = oracle.Extractor(1000, 1.2, 8, 20, 7)
    base, _ = ox.extract(img)
    existing = base[::7].copy()          # pretend these were direct-tracked from the last frame
    okps, odesc = ox.extract_frame(img, existing)
    kps, desc = ex.extract_frame(img, existing)
    assert_kps_equal(kps, okps, "extract_frame")
    assert (desc == odesc).all()


def test_async_two_handles_equals_sync(gpu, synth):
    """ygz_orb_extract_async / ygz_orb_wait with two handles in flight returns what the blocking call returns."""
    import ctypes as C
    import numpy as np
    w, h, B = 320, 240, 3
    imgs = np.stack([synth.synth_frame(40 + i, w, h) for i in range(2 * B)])
    exs = [gpu.ORBextractor(500, 1.2, 8, 20, 7, width=w, height=h, max_batch=B) for _ in range(2)]
    ref = exs[0].extract_batch(imgs[:B]) + exs[0].extract_batch(imgs[B:])
    cap = exs[0].cap
    bufs = []
    for k, ex in enumerate(exs):
        kps = np.zeros((B, cap), gpu.KP_DTYPE)
        desc = np.zeros((B, cap, 32), np.uint8)
        cnt = np.zeros(B, np.int32)
        part = np.ascontiguousarray(imgs[k * B:(k + 1) * B])
        gpu._check(gpu.lib().ygz_orb_extract_async(ex.h, part.ctypes.data, B, w * h, w, kps.ctypes.data, desc.ctypes.data,
                                                   cnt.ctypes.data, cap), "ygz_orb_extract_async")
        bufs.append((kps, desc, cnt, part))
    for ex in exs:
        gpu._check(gpu.lib().ygz_orb_wait(ex.h), "ygz_orb_wait")
    got = []
    for kps, desc, cnt, _ in bufs:
        got += [(kps[f, :cnt[f]], desc[f, :cnt[f]]) for f in range(B)]
    for (k0, d0), (k1, d1) in zip(ref, got):
        assert_kps_equal(k0, k1, "async")
        assert (d0 == d1).all()
