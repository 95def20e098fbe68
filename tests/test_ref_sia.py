"""The oracle's SparseImgAlign restatement (oracle/sparse_align.cpp) against the REFERENCE's own ygz::SparseImgAlign:
src/SparseImageAlign.cc + include/NLSSolver_impl.hpp compiled unmodified against oracle/ref_mock (oracle/_ref/libsia_ref.so,
oracle/ref_sia_shim.cpp).  Pins everything in the reference tree (level loop, reference-patch / Jacobian cache, residual and
H / Jres accumulation, Gauss-Newton accept / rollback / stop logic); Eigen's LDLT and Sophus' SE3 are absent third-party code whose
stand-ins forward to the oracle's restatement, so they are common to both sides (stated in oracle/ref_sia_shim.cpp).
Result: pose, H and chi2 BIT-identical, tracked counts equal — on well-conditioned problems, bad initial guesses (rollback path),
partially invalid feature sets and the degenerate cases.  Build container only."""
import numpy as np
import pytest

import ref_lib

W, H = 752, 480


def build(pkg, oracle, synth, seed, n_pts=200, nlevels=4, init=None, valid=None, w=W, h=H):
    s = synth.synth_sia_pair(seed, n_pts=n_pts, w=w, h=h)
    ox = oracle.Extractor(100, 2.0, nlevels, 20, 7)
    ox.extract(s["ref"]); ref = [ox.level(l) for l in range(nlevels)]
    ox.extract(s["cur"]); cur = [ox.level(l) for l in range(nlevels)]
    inv = ox.tables()["inv_scale"]
    cur_pose = init if init is not None else ((0, 0, 0, 1), (0, 0, 0))
    return pkg.make_sia_problem(ref, cur, inv, s["K"], s["px"], s["xyz"], valid=valid, cur_pose=cur_pose, cls=oracle.SiaProblem), s


def same(o, r):
    on, oq, ot, oH, otr = o
    rn, rq, rt, rH, rchi2 = r
    assert on == rn, (on, rn)
    assert (oq.view(np.uint32) == rq.view(np.uint32)).all() and (ot.view(np.uint32) == rt.view(np.uint32)).all(), (oq, rq, ot, rt)
    assert (oH.view(np.uint32) == rH.view(np.uint32)).all()
    if len(otr):
        accepted = otr[otr[:, 5] == 1]
        if len(accepted):
            assert np.float32(accepted[-1, 2]).view(np.uint32) == np.float32(rchi2).view(np.uint32)     # chi2_ of the last accepted iteration


@pytest.fixture(scope="module")
def have_ref():
    if ref_lib._load("libsia_ref.so") is None:
        pytest.skip("oracle/_ref/libsia_ref.so not built (no /root/reference here)")


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 5])
def test_run_matches_reference(have_ref, pkg, oracle, synth, seed):
    p, s = build(pkg, oracle, synth, seed)
    o = oracle.sia_run(p, 3, 1, 10, 1e-6)
    same(o, ref_lib.sia_run(p, 3, 1, 10))
    assert o[0] > 150


def test_rollback_and_partial_validity(have_ref, pkg, oracle, synth):
    init = ((0.0, 0.0, 0.01, 0.99995), (0.15, -0.1, 0.05))           # far-off initial guess
    p, _ = build(pkg, oracle, synth, 12, init=init)
    same(oracle.sia_run(p, 3, 1, 10, 1e-6), ref_lib.sia_run(p, 3, 1, 10))
    # Rollback path (new_chi2 > 1.2 * chi2_ -> model = old_model; break, NLSSolver_impl.hpp:53-63): Gauss-Newton on the rendered pairs
    # never increases chi2 that much, so the current image is noise and only 2-5 patches are valid (erratic steps)
    rejected = 0
    for seed in range(30, 40):
        s = synth.synth_sia_pair(seed, n_pts=200, w=W, h=H)
        rng = np.random.default_rng(seed)
        ox = oracle.Extractor(100, 2.0, 4, 20, 7)
        ox.extract(s["ref"]); ref = [ox.level(l) for l in range(4)]
        ox.extract(rng.integers(0, 256, (H, W), dtype=np.uint8)); cur = [ox.level(l) for l in range(4)]
        valid = np.zeros(200, np.uint8); valid[rng.choice(200, 2 + seed % 4, replace=False)] = 1
        p = pkg.make_sia_problem(ref, cur, ox.tables()["inv_scale"], s["K"], s["px"], s["xyz"], valid=valid, cls=oracle.SiaProblem)
        o = oracle.sia_run(p, 3, 1, 10, 1e-6)
        same(o, ref_lib.sia_run(p, 3, 1, 10))
        rejected += int((o[4][:, 5] == 0).sum())
    assert rejected >= 5                                              # the rollback path was taken
    rng = np.random.default_rng(3)
    valid = (rng.random(200) < 0.6).astype(np.uint8)                  # null / bad map points and outliers
    p, _ = build(pkg, oracle, synth, 4, valid=valid)
    same(oracle.sia_run(p, 3, 1, 10, 1e-6), ref_lib.sia_run(p, 3, 1, 10))
    p, _ = build(pkg, oracle, synth, 6)
    same(oracle.sia_run(p, 2, 0, 10, 1e-6), ref_lib.sia_run(p, 2, 0, 10))       # other level ranges, down to the full-resolution level
    # upstream quirk: run() overwrites n_iter_ with its own table {10, 10, ...} per level (src/SparseImageAlign.cc:40-45), so the
    # constructor's n_iter is ignored — the reference asked for 4 iterations behaves like the oracle asked for 10
    same(oracle.sia_run(p, 3, 3, 10, 1e-6), ref_lib.sia_run(p, 3, 3, 4))
