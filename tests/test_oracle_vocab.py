"""CPU checks of the oracle's DBoW2 vocabulary restatement (oracle/vocabulary.cpp).  Where the reference tree is present
(build container) the loader is pinned on the reference's own Vocabulary/ORBvoc.bin: header, node / word counts, depth
and the k^2 = 100 distinct level-2 nodes that Frame::ComputeBoW's levelsup = 4 yields."""
import os
import numpy as np
import pytest

VOC = "/root/reference/Vocabulary/ORBvoc.bin"


def test_synthetic_vocabulary_roundtrip(oracle, synth):
    data = synth.synth_vocabulary(3, k=6, L=3)
    v = oracle.Vocabulary(data=data)
    assert (v.k, v.L, v.scoring, v.weighting) == (6, 3, 0, 0)
    assert v.n_nodes == (len(data) - 24) // 41 + 1 and 0 < v.n_words < v.n_nodes
    rng = np.random.default_rng(0)
    desc = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    wid, w, nid = v.transform_features(desc, levelsup=2)          # node at level 1: one of the k children of the root
    assert set(np.unique(nid)) <= set(range(1, 7)) and (wid >= 0).all() and (wid < v.n_words).all()
    ids, vals, fv = v.transform(desc, levelsup=2)
    assert (np.diff(ids) > 0).all() and abs(vals.sum() - 1.0) < 1e-12
    kept = sorted(i for l in fv.values() for i in l)
    assert kept == list(np.nonzero(w > 0)[0])                    # stopped words (weight 0) are in neither vector
    assert all(l == sorted(l) for l in fv.values())
    wid0, _, nid0 = v.transform_features(desc, levelsup=5)       # levelsup >= L: node = root
    assert (nid0 == 0).all() and (wid0 == wid).all()


def test_reference_orbvoc(oracle):
    if not os.path.exists(VOC):
        pytest.skip("reference vocabulary not present on this box")
    v = oracle.Vocabulary(path=VOC)
    assert (v.k, v.L, v.scoring, v.weighting) == (10, 6, 0, 0)
    assert v.n_nodes == 1082074 and v.n_words == 971815           # 24-byte header + 1 082 073 records of 41 B = 44 365 017 B
    rng = np.random.default_rng(1)
    desc = rng.integers(0, 256, (4000, 32), dtype=np.uint8)
    wid, w, nid = v.transform_features(desc, levelsup=4)
    assert len(np.unique(nid)) == 100 and (w >= 0).all() and (w > 0).mean() > 0.9 and wid.max() < v.n_words
