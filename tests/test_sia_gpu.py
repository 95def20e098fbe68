"""GPU parity of SparseImgAlign::run (through the C ABI) against the oracle.
Tolerance (BASELINE north_star): SE(3) pose within 1e-4 of the CPU path, identical tracked count."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
POSE_TOL = 1e-4


def build_problem(gpu, oracle, synth, seed, cls, n_pts=200, w=752, h=480, nlevels=4, init=None, valid=None, **kw):
    s = synth.synth_sia_pair(seed, n_pts=n_pts, w=w, h=h, **kw)
    ox = oracle.Extractor(100, 2.0, nlevels, 20, 7)           # Examples/Monocular/EuRoC.yaml:36-39 (scale 2.0, 4 levels)
    ox.extract(s["ref"])
    ref_levels = [ox.level(l) for l in range(nlevels)]
    ox.extract(s["cur"])
    cur_levels = [ox.level(l) for l in range(nlevels)]
    inv_scale = ox.tables()["inv_scale"]
    cur_pose = init if init is not None else ((0, 0, 0, 1), (0, 0, 0))
    pr = gpu.make_sia_problem(ref_levels, cur_levels, inv_scale, s["K"], s["px"], s["xyz"], valid=valid, cur_pose=cur_pose, cls=cls)
    return pr, s


def pose_diff(q1, t1, q2, t2):
    dq = min(np.abs(q1 - q2).max(), np.abs(q1 + q2).max())
    return max(float(dq), float(np.abs(t1 - t2).max()))


def assert_traces_equivalent(tr, otr, eps=1e-6, chi2_rtol=1e-3):
    """Per level: the same iteration numbers, measurement counts, accept / reject decisions and chi2 (rtol 1e-3) on both sides.
    The early stop `max|x| <= eps_` (include/NLSSolver_impl.hpp:87-89, eps_ = 1e-6) is decided at the rounding-noise floor of the update
    once a level has converged — float sums taken in a different order move max|x| by a few 1e-7 — so a level may end one iteration
    earlier on one side; the surplus records of the other side must then be accepted updates below 4 eps."""
    levels = sorted(set(tr[:, 0].tolist()) | set(otr[:, 0].tolist()), reverse=True)
    for lvl in levels:
        a, b = tr[tr[:, 0] == lvl], otr[otr[:, 0] == lvl]
        n = min(len(a), len(b))
        assert n >= 1, "level %d missing on one side" % lvl
        assert (a[:n, 1] == b[:n, 1]).all() and (a[:n, 3] == b[:n, 3]).all() and (a[:n, 5] == b[:n, 5]).all(), (lvl, a[:n], b[:n])
        if chi2_rtol is not None:
            assert np.allclose(a[:n, 2], b[:n, 2], rtol=chi2_rtol, atol=1e-3)
        if len(a) != len(b):
            shorter, longer = (a, b) if len(a) < len(b) else (b, a)
            extra = longer[n:]
            assert shorter[-1, 5] == 1 and shorter[-1, 4] <= eps, "the shorter side did not stop on eps: %r" % (shorter[-1],)
            assert len(extra) <= 2 and (extra[:, 5] == 1).all() and (extra[:, 4] <= 4 * eps).all() and longer[n - 1, 4] <= 4 * eps, extra


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_sia_pose_matches_oracle(gpu, oracle, synth, seed):
    pg, s = build_problem(gpu, oracle, synth, seed, gpu.SiaProblem)
    po, _ = build_problem(gpu, oracle, synth, seed, oracle.SiaProblem)
    sia = gpu.SparseImgAlign(3, 1, 10)                         # SparseImgAlign(nLevels-1, 1), src/Tracking.cc:207
    n, q, t, Hm = sia.run(pg)
    on, oq, ot, oH, otr = oracle.sia_run(po, 3, 1, 10, 1e-6)
    assert n == on and n > 150
    assert pose_diff(q, t, oq, ot) <= POSE_TOL, (q, t, oq, ot)
    # both must actually have converged to the rendered motion
    assert pose_diff(q, t, s["q_true"], s["t_true"]) < 5e-3
    assert_traces_equivalent(sia.trace(), otr)
    assert np.allclose(Hm, oH, rtol=2e-3, atol=1e-2)


def test_sia_batch_and_edge_cases(gpu, oracle, synth):
    seeds = [4, 5, 6]
    sia = gpu.SparseImgAlign(3, 1, 10, max_problems=4)
    pgs, pos = [], []
    rng = np.random.default_rng(0)
    for sd in seeds:
        valid = (rng.random(200) < 0.8).astype(np.uint8)       # null / bad / outlier map points
        pgs.append(build_problem(gpu, oracle, synth, sd, gpu.SiaProblem, valid=valid)[0])
        pos.append(build_problem(gpu, oracle, synth, sd, oracle.SiaProblem, valid=valid)[0])
    # a problem whose features all sit on the image border: nothing visible, T unchanged, returns 0
    s = synth.synth_sia_pair(9)
    px = np.full((50, 2), 1.0, np.float32)
    edge_kw = dict(valid=None)
    pg_e, _ = build_problem(gpu, oracle, synth, 9, gpu.SiaProblem)
    po_e, _ = build_problem(gpu, oracle, synth, 9, oracle.SiaProblem)
    for p in (pg_e, po_e):
        p.refs[5][:] = 1.0                                     # ref_px array -> (1,1): outside every 3-px border
    pgs.append(pg_e)
    pos.append(po_e)
    q, t, Hm, nt = sia.run_batch(pgs)
    for i, po in enumerate(pos):
        on, oq, ot, oH, _ = oracle.sia_run(po, 3, 1, 10, 1e-6)
        assert nt[i] == on, (i, nt[i], on)
        assert pose_diff(q[i], t[i], oq, ot) <= POSE_TOL
    assert nt[3] == 0


def test_sia_rollback_from_bad_init(gpu, oracle, synth):
    """Far-off initial guess: exercises the chi2 rollback / early-stop paths; GPU must follow the oracle.

    On this ill-conditioned problem the reference's OWN result is only determined to a few 1e-4: its float, sequential
    accumulation of H / Jres (src/SparseImageAlign.cc:216-217) is ~1e-6 relative away from the exact sums, and the bad
    conditioning of the first iterations amplifies that (the exact-sums run may even take a different accept / reject path).  The
    device sums per-thread partials in float and reduces them in double — neither the reference's order nor exact.
    The bound tested here is therefore max(1e-4, the CPU path's own float-vs-exact-sums spread, measured in
    the test), with the same accept / reject decisions required on top (assert_traces_equivalent); the seven well-conditioned seeds keep the plain 1e-4."""
    init = ((0.0, 0.0, 0.01, 0.99995), (0.15, -0.1, 0.05))
    pg, _ = build_problem(gpu, oracle, synth, 12, gpu.SiaProblem, init=init)
    po, _ = build_problem(gpu, oracle, synth, 12, oracle.SiaProblem, init=init)
    sia = gpu.SparseImgAlign(3, 1, 10)
    n, q, t, _ = sia.run(pg)
    on, oq, ot, _, otr = oracle.sia_run(po, 3, 1, 10, 1e-6)
    assert_traces_equivalent(sia.trace(), otr, chi2_rtol=None)     # accept / reject decisions, iteration for iteration
    assert n == on
    # how well is the CPU result itself determined?  Same algorithm with exact (double) sums:
    oracle.sia_wide_sums(True)
    try:
        wn, wq, wt, _, wtr = oracle.sia_run(po, 3, 1, 10, 1e-6)
    finally:
        oracle.sia_wide_sums(False)
    d_gpu = pose_diff(q, t, oq, ot)                   # GPU vs the reference's arithmetic (float, sequential sums)
    d_cpu = pose_diff(oq, ot, wq, wt)                 # the reference's float sums vs exact sums: the CPU path's own indeterminacy
    print("bad-init case: |gpu - cpu| = %.2e, |cpu(float sums) - cpu(exact sums)| = %.2e, traces %d / %d iterations" %
          (d_gpu, d_cpu, len(otr), len(wtr)))
    assert d_gpu <= max(POSE_TOL, d_cpu), (d_gpu, d_cpu)


def test_sia_run_orb_reads_the_extractor_pyramids_in_place(gpu, oracle, synth):
    """ygz_sia_run_orb (level images taken from ORBextractor handles, pitch != width) == ygz_sia_run on the downloaded levels, bit for
    bit (same kernel, same pixels), and == the oracle within the pose tolerance.  One handle holds [ref0, cur0, ref1, cur1]; a second
    single-frame handle provides a current frame from ANOTHER handle / stream."""
    nl, W, Hh = 4, 752, 480
    scenes = [synth.synth_sia_pair(sd, n_pts=200, w=W, h=Hh) for sd in (11, 12)]
    ex = gpu.ORBextractor(100, 2.0, nl, 20, 7, width=W, height=Hh, max_batch=4)
    ex2 = gpu.ORBextractor(100, 2.0, nl, 20, 7, width=W, height=Hh, max_batch=1)
    ex.extract_batch([scenes[0]["ref"], scenes[0]["cur"], scenes[1]["ref"], scenes[1]["cur"]])
    ex2.extract_batch([scenes[1]["cur"]])
    inv = ex.tables()["inv_scale"]
    rng = np.random.default_rng(3)
    valid = [(rng.random(200) < 0.85).astype(np.uint8) for _ in scenes]
    feats = [gpu.make_sia_features(s["K"], s["px"], s["xyz"], valid=v) for s, v in zip(scenes, valid)]
    sia = gpu.SparseImgAlign(3, 1, 10, max_problems=4, max_features=256)
    q, t, Hm, nt = sia.run_batch_orb(ex, ex, feats, [0, 2], [1, 3])
    traces = [sia.trace(i) for i in range(2)]
    # the same through host level images
    host = []
    for i, (s, v) in enumerate(zip(scenes, valid)):
        ref = [ex.level(2 * i, l) for l in range(nl)]
        cur = [ex.level(2 * i + 1, l) for l in range(nl)]
        host.append(gpu.make_sia_problem(ref, cur, inv, s["K"], s["px"], s["xyz"], valid=v))
    q2, t2, H2, nt2 = sia.run_batch(host)
    assert (nt == nt2).all() and (nt > 120).all()
    assert np.array_equal(q, q2) and np.array_equal(t, t2) and np.array_equal(Hm, H2)
    for i in range(2):
        assert np.array_equal(traces[i], sia.trace(i))
    # current frame from a second handle (its own stream): same answer as problem 1
    q3, t3, H3, nt3 = sia.run_batch_orb(ex, ex2, [feats[1]], [2], [0])
    assert nt3[0] == nt[1] and np.array_equal(q3[0], q[1]) and np.array_equal(t3[0], t[1])
    # oracle
    for i, (s, v) in enumerate(zip(scenes, valid)):
        po, _ = build_problem(gpu, oracle, synth, (11, 12)[i], oracle.SiaProblem, valid=v)
        on, oq, ot, oH, _ = oracle.sia_run(po, 3, 1, 10, 1e-6)
        assert nt[i] == on
        assert pose_diff(q[i], t[i], oq, ot) <= POSE_TOL
    # argument errors: frame index outside the handle, level beyond the pyramid
    with pytest.raises(Exception):
        sia.run_batch_orb(ex, ex, [feats[0]], [0], [4])
    deep = gpu.SparseImgAlign(4, 1, 10, max_problems=1, max_features=256)
    with pytest.raises(Exception):
        deep.run_batch_orb(ex, ex, [feats[0]], [0], [1])


def test_sia_many_problems_per_launch(gpu, oracle, synth):
    """320 problems in one launch (more CTAs than the 2 x 148 the chip holds at once; masked features; per-problem traces): every problem
    within the oracle bound, and bit-equal to the same problem solved in a small batch."""
    seeds = [0, 1, 2, 4, 5, 6]
    rng = np.random.default_rng(1)
    pgs, pos = [], []
    for sd in seeds:
        valid = (rng.random(200) < 0.85).astype(np.uint8) if sd % 2 else None
        pgs.append(build_problem(gpu, oracle, synth, sd, gpu.SiaProblem, valid=valid)[0])
        pos.append(build_problem(gpu, oracle, synth, sd, oracle.SiaProblem, valid=valid)[0])
    ref = [oracle.sia_run(po, 3, 1, 10, 1e-6) for po in pos]
    P = 320
    sia = gpu.SparseImgAlign(3, 1, 10, max_problems=P)
    q, t, Hm, nt = sia.run_batch([pgs[i % len(seeds)] for i in range(P)])
    for i in range(P):
        on, oq, ot, oH, otr = ref[i % len(seeds)]
        assert nt[i] == on, (i, nt[i], on)
        assert pose_diff(q[i], t[i], oq, ot) <= POSE_TOL, i
        assert np.allclose(Hm[i].reshape(6, 6), oH, rtol=2e-3, atol=1e-2)
    for i in (0, 3, 319):
        assert_traces_equivalent(sia.trace(i), ref[i % len(seeds)][4])
    # a problem's result does not depend on the batch it is solved in
    small = gpu.SparseImgAlign(3, 1, 10, max_problems=8)
    q2, t2, _, nt2 = small.run_batch(pgs)
    for i in range(len(seeds)):
        assert nt2[i] == nt[i] and np.array_equal(q2[i], q[i]) and np.array_equal(t2[i], t[i])
