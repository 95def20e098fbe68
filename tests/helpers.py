"""Shared builders for matcher / aligner tests (views over numpy arrays for both the oracle and the product)."""
import numpy as np

SCALE8 = (np.float32(1.2) ** 0 * np.cumprod(np.concatenate([[1.0], np.full(7, 1.2)]))).astype(np.float32)


def hamming_matrix(a, b):
    x = np.bitwise_xor(a[:, None, :], b[None, :, :])
    return np.unpackbits(x, axis=2).sum(2).astype(np.int32)


def synth_vocab_nodes(desc, n_nodes=48, seed=5):
    """Synthetic BoW assignment: node = nearest of n_nodes random centroid descriptors (stands in for
    DBoW2 transform(levelsup=4); the FeatureVector is an INPUT format of SearchByBoW)."""
    rng = np.random.default_rng(seed)
    cent = rng.integers(0, 256, (n_nodes, 32), dtype=np.uint8)
    if len(desc) == 0:
        return np.zeros(0, np.uint32)
    d = hamming_matrix(np.ascontiguousarray(desc), cent)
    return (d.argmin(1) * 7 + 3).astype(np.uint32)      # arbitrary, non-contiguous node ids


def two_frames(oracle, synth, seed=1, shift=(9, 5), w=752, h=480, nfeatures=1000):
    """Two overlapping crops of one texture -> oracle keypoints/descriptors of both."""
    tex = synth.synth_frame(seed, w + 64, h + 64)
    a = np.ascontiguousarray(tex[16:16 + h, 16:16 + w])
    b = np.ascontiguousarray(tex[16 + shift[1]:16 + shift[1] + h, 16 + shift[0]:16 + shift[0] + w])
    ox = oracle.Extractor(nfeatures, 1.2, 8, 20, 7)
    ka, da = ox.extract(a)
    kb, db = ox.extract(b)
    return (a, ka, da), (b, kb, db)
