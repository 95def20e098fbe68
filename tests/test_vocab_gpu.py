"""GPU parity of the DBoW2 vocabulary transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1151-1292) against the
oracle: word id, weight and level-(L - levelsup) node id per descriptor identical; BowVector (float64, L1-normalised) and
FeatureVector identical; zero-copy transform of an extractor's device descriptors feeds the batched SearchByBoW."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,L,levelsup", [(10, 3, 1), (10, 4, 2), (6, 3, 2), (3, 5, 4), (10, 2, 4)])
def test_transform_bit_exact(gpu, oracle, synth, k, L, levelsup):
    data = synth.synth_vocabulary(11 + k + L, k=k, L=L)
    gv = gpu.ORBVocabulary(data=data)
    ov = oracle.Vocabulary(data=data)
    assert (gv.k, gv.L, gv.n_nodes, gv.n_words) == (ov.k, ov.L, ov.n_nodes, ov.n_words)
    q, t, _ = synth.synth_descriptors(5, 1500, 1500)
    desc = np.concatenate([q, t])
    wid, w, nid = gv.transform_features(desc, levelsup)
    owid, ow, onid = ov.transform_features(desc, levelsup)
    assert (wid == owid).all() and (nid == onid).all() and (w.view(np.uint32) == ow.view(np.uint32)).all()
    ids, vals, fv = gv.transform(desc[:1000], levelsup)
    oids, ovals, ofv = ov.transform(desc[:1000], levelsup)
    assert (ids == oids).all() and (vals == ovals).all() and fv == ofv


def test_transform_orb_feeds_search_by_bow(gpu, oracle, synth):
    """Frame::ComputeBoW on the device for a batch of frames, then the batched SearchByBoW with those node ids equals the
    oracle's SearchByBoW on FeatureVectors built from the oracle's transform."""
    import helpers as H
    w, h, B = 480, 360, 3
    frames = synth.synth_stream(21, B, w, h)
    data = synth.synth_vocabulary(4, k=10, L=4)
    gv = gpu.ORBVocabulary(data=data)
    ov = oracle.Vocabulary(data=data)
    ex = gpu.ORBextractor(600, 1.2, 8, 20, 7, width=w, height=h, max_batch=B)
    res = ex.extract_batch(frames)
    out = gv.transform_orb(ex, B, levelsup=2, want=("node", "word", "weight"), timed=True)
    for f, (kps, desc) in enumerate(res):
        owid, ow, onid = ov.transform_features(desc, 2)
        n = len(kps)
        assert (out["node"][f, :n] == onid).all() and (out["word"][f, :n] == owid).all()
        assert (out["weight"][f, :n].view(np.uint32) == ow.view(np.uint32)).all()
        assert (out["node"][f, n:] == -1).all()
    mt = gpu.ORBmatcher(0.7, True)
    kf_idx = np.array([0, 1], np.int32); f_idx = np.array([1, 2], np.int32)
    node = out["node"].astype(np.uint32)
    nm, matches = mt.SearchByBoWBatch(ex, kf_idx, f_idx, node)
    for p in range(2):
        (ka, da), (kb, db) = res[kf_idx[p]], res[f_idx[p]]
        mp = np.arange(len(ka), dtype=np.int32)
        kf = gpu.make_frame_view(ka, da, map_point=mp, node_of_feature=out["node"][kf_idx[p], :len(ka)].astype(np.uint32), cls=oracle.FrameView)
        fr = gpu.make_frame_view(kb, db, node_of_feature=out["node"][f_idx[p], :len(kb)].astype(np.uint32), cls=oracle.FrameView)
        mps = gpu.make_map_points(np.zeros((len(ka), 3)), da, cls=oracle.MapPoints)
        o = np.zeros(len(kb), np.int32)
        n_bow = oracle.lib().orc_search_by_bow(0.7, 1, kf.ptr(), mps.ptr(), fr.ptr(), o.ctypes.data_as(oracle.i32p))
        assert nm[p] == n_bow and (matches[p, :len(kb)] == o).all()


def test_transform_real_vocabulary_shape(gpu, oracle, synth):
    """k = 10, L = 6: the shape of the reference's ORBvoc.bin (1 082 074 nodes there, ~0.98 M here; 32 MB of node descriptors,
    larger than L1 and than any synthetic tree of the other tests) — word, weight and level-2 node (levelsup = 4) identical."""
    data = synth.synth_vocabulary(2, k=10, L=6)
    gv = gpu.ORBVocabulary(data=data)
    ov = oracle.Vocabulary(data=data)
    assert (gv.k, gv.L) == (10, 6) and gv.n_nodes == ov.n_nodes > 900000 and gv.n_words == ov.n_words
    q, t, _ = synth.synth_descriptors(9, 2000, 2000)
    desc = np.concatenate([q, t])
    wid, w, nid = gv.transform_features(desc, 4)
    owid, ow, onid = ov.transform_features(desc, 4)
    assert (wid == owid).all() and (nid == onid).all() and (w.view(np.uint32) == ow.view(np.uint32)).all()
    assert len(np.unique(nid)) > 50                    # level-2 nodes: up to 100 buckets, as SearchByBoW sees them


def test_attached_vocabulary_equals_explicit_nodes(gpu, synth):
    """ygz_matcher_set_vocabulary: Frame::ComputeBoW inside the batched SearchByBoW (node ids never leave the device) gives the
    matches of the same search fed with the node ids of ygz_vocab_transform_orb; sync and async entry points."""
    import ctypes as C
    w, h, B = 480, 360, 4
    frames = synth.synth_stream(23, B, w, h)
    gv = gpu.ORBVocabulary(data=synth.synth_vocabulary(6, k=10, L=4))
    ex = gpu.ORBextractor(600, 1.2, 8, 20, 7, width=w, height=h, max_batch=B)
    ex.extract_batch(frames)
    tr = gv.transform_orb(ex, B, levelsup=2, want=("node", "weight"))
    node = np.where(tr["weight"] > 0, tr["node"], -1).astype(np.uint32)     # stopped words are not in the FeatureVector
    assert (tr["weight"] == 0).sum() > (tr["node"] >= 0).sum() * 0.01       # the synthetic vocabulary stops ~5 % of its words
    kf_idx = np.array([3, 0, 1, 2], np.int32); f_idx = np.arange(4, dtype=np.int32)
    mt = gpu.ORBmatcher(0.7, True)
    nm0, m0 = mt.SearchByBoWBatch(ex, kf_idx, f_idx, node)
    nm_syn, _ = mt.SearchByBoWBatch(ex, kf_idx, f_idx, None)             # synthetic stand-in: different buckets
    mt.set_vocabulary(gv, 2)
    nm1, m1 = mt.SearchByBoWBatch(ex, kf_idx, f_idx, None)
    assert (nm1 == nm0).all() and (m1 == m0).all() and nm0.sum() > 100 and (nm_syn != nm0).any()
    cap = ex.cap
    out_m = np.zeros((4, cap), np.int32); out_n = np.zeros(4, np.int32)
    gpu._check(gpu.lib().ygz_search_by_bow_batch_async(mt.h, ex.h, C.c_float(0.7), 1, kf_idx.ctypes.data_as(gpu.i32p),
                                                       f_idx.ctypes.data_as(gpu.i32p), 4, out_m.ctypes.data_as(gpu.i32p),
                                                       out_n.ctypes.data_as(gpu.i32p)), "async")
    gpu._check(gpu.lib().ygz_orb_wait(ex.h), "wait")
    assert (out_n == nm0).all() and (out_m == m0).all()
    mt.set_vocabulary(None)
    nm2, _ = mt.SearchByBoWBatch(ex, kf_idx, f_idx, None)
    assert (nm2 == nm_syn).all()


def test_attached_vocabulary_equals_oracle_frontend(gpu, oracle, synth):
    """The bench step end to end against the CPU arm: extractor + ComputeBoW (stopped words excluded) + SearchByBoW vs the
    previous frame — total matches and key points of the device path == oracle.frontend_batch with the same vocabulary."""
    w, h, B = 480, 360, 6
    frames = synth.synth_stream(29, B, w, h)
    data = synth.synth_vocabulary(8, k=10, L=4)
    gv = gpu.ORBVocabulary(data=data)
    ex = gpu.ORBextractor(600, 1.2, 8, 20, 7, width=w, height=h, max_batch=B)
    res = ex.extract_batch(frames)
    mt = gpu.ORBmatcher(0.7, True)
    mt.set_vocabulary(gv, 2)
    kf_idx = np.array([(i - 1) % B for i in range(B)], np.int32); f_idx = np.arange(B, dtype=np.int32)
    nm, _ = mt.SearchByBoWBatch(ex, kf_idx, f_idx, None)
    om, ok = oracle.frontend_batch(frames, 1, nfeatures=600, native=False, vocab=oracle.Vocabulary(data=data), levelsup=2)
    assert ok == sum(len(k) for k, _ in res) and om == int(nm.sum())
