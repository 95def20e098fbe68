"""bench.py --workload c5 | c3 — the two other GPU configurations of BASELINE.json as first-class bench lines (same JSON contract
as the C2 headline: value / e2e / roofline / cpu_baseline / clocks / gpu_launches; `--impl reference` runs the CPU arm alone).

C5  BASELINE.json configs[4]: stereo 1242x375 (KITTI shape, synthetic), nfeatures 2000, 8 levels x 1.2: one step = B stereo pairs
    through ORBextractor x 2 (left and right handle, each on its own stream) + Frame::ComputeStereoMatches for every pair (row-band
    Hamming association, 11x11 SAD refinement, median cut: src/Frame.cc:509-682) + the dense brute-force matcher (best / second
    best right descriptor of every left descriptor), all on the extractors' device-resident outputs.  Unit: stereo pairs/s.
C3  configs[2]: SparseImgAlign, 4-level x2.0 pyramid, 200 patches, levels 3..1, 10 Gauss-Newton iterations per level
    (src/SparseImageAlign.cc:20-244, src/Tracking.cc:207): one step = P independent alignments in one launch.  Unit: alignments/s.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
C5_W, C5_H = 1242, 375
C5_ORB = dict(nfeatures=2000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7)
C5_MB, C5_MBF = 0.54, 386.1448                  # Examples/Stereo/KITTI00-02.yaml: bf = 386.1448, baseline 0.54 m
C3_W, C3_H = 752, 480
C3_LEVELS, C3_MAX, C3_MIN, C3_ITERS, C3_PATCHES = 4, 3, 1, 10, 200


def _synth():
    import __graft_entry__ as g
    g.load_package()
    import importlib
    return importlib.import_module("orb_ygz_slam_b200.synth")


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    try:
        O.lib(native=True); native = True
    except Exception:
        native = False
    return O, native


def _level_sizes(w, h, nlevels, scale):
    out, s = [], np.float32(1.0)
    for l in range(nlevels):
        inv = np.float32(1.0) / s
        out.append((int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))))
        s = np.float32(s * np.float32(scale))
    return out


# ----------------------------------------------------------------------------------------------------------------- C5
def c5_pairs(seed, n):
    synth = _synth()
    L = np.empty((n, C5_H, C5_W), np.uint8); R = np.empty((n, C5_H, C5_W), np.uint8)
    for i in range(n):
        l, r, _, _ = synth.synth_stereo_pair(seed * 1000 + i, C5_W, C5_H)
        L[i], R[i] = l, r
    return L, R


def c5_config(batch):
    return {"workload": "C5 synthetic KITTI-shape stereo 1242x375 u8 x2: ORBextractor(2000,1.2,8,20,7) on left and right + "
                        "Frame::ComputeStereoMatches (row-band Hamming, 11x11 SAD refinement, median cut) + dense brute-force matcher "
                        "(best / second best of ~2000 x ~2000 descriptors) per pair",
            "pairs_per_step": batch,
            "l2_policy": "inputs larger than L2 (%d distinct pairs/step, pyramid working set %.0f MB)" % (batch, 2 * batch * 1.44),
            "parallelism": "replicas: one independent batch of pairs per GPU"}


def c5_cpu_arm(L, R, threads, min_seconds, check=None):
    """Oracle port: per pair extract L, extract R, ComputeStereoMatches, dense best/second-best; one pair per thread.
    check = dict(u_right=..., n_matches=...) of the GPU arm for pair 0: parity gate (bit-exact) inside the bench."""
    from concurrent.futures import ThreadPoolExecutor
    import threading
    O, native = _oracle()
    import __graft_entry__ as g
    m = g.load_package()
    local = threading.local()

    def one(i):
        if not hasattr(local, "ex"):
            kw = dict(nfeatures=C5_ORB["nfeatures"], scale=C5_ORB["scaleFactor"], nlevels=C5_ORB["nlevels"], ini_th=C5_ORB["iniThFAST"],
                      min_th=C5_ORB["minThFAST"], native=native)
            local.ex, local.ex2 = O.Extractor(**kw), O.Extractor(**kw)
            local.t = local.ex.tables()
        kL, dL = local.ex.extract(L[i])
        pyrL = m.make_direct_frame([local.ex.level(l) for l in range(8)])
        kR, dR = local.ex2.extract(R[i])
        pyrR = m.make_direct_frame([local.ex2.level(l) for l in range(8)])
        u, z, n = O.compute_stereo_matches(pyrL, pyrR, kL, dL, kR, dR, local.t["scale"], local.t["inv_scale"], C5_MB, C5_MBF)
        bi, bd, sd = O.bf_top2(dL, dR)
        return n, u, len(kL)

    n = len(L)
    with ThreadPoolExecutor(threads) as pool:
        first = list(pool.map(one, range(min(n, threads))))
        gate = None
        if check is not None:
            n0, u0, nk = first[0]
            gate = bool(n0 == int(check["n_matches"][0]) and (u0.view(np.uint32) == check["u_right"][0, :nk].view(np.uint32)).all())
        done, t0 = 0, time.perf_counter()
        while True:
            list(pool.map(one, range(n)))
            done += n
            dt = time.perf_counter() - t0
            if dt >= min_seconds:
                break
    return done / dt, done, dt, native, gate


def run_c5(args, B, ctx):
    """ctx: dict(m, torch, dist, rank, world, local, ClockSampler, host_threads, measured_peak_gbs, emit_json, shard)."""
    m, torch, dist, rank, world, local = ctx["m"], ctx["torch"], ctx["dist"], ctx["rank"], ctx["world"], ctx["local"]
    Limg, Rimg = c5_pairs(1 + rank, B)
    npairs = 2                                               # handle pairs in flight on the host-buffer path
    hs = []
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
    for _ in range(npairs):
        exL = m.ORBextractor(width=C5_W, height=C5_H, max_batch=B, device=local, **C5_ORB)
        exR = m.ORBextractor(width=C5_W, height=C5_H, max_batch=B, device=local, **C5_ORB)
        cap = exL.cap
        b = dict(L=pin((B, C5_H, C5_W), torch.uint8), R=pin((B, C5_H, C5_W), torch.uint8))
        b["L"].numpy()[:] = Limg; b["R"].numpy()[:] = Rimg
        for s in ("L", "R"):
            b["kps" + s] = pin((B, cap, 7), torch.float32); b["desc" + s] = pin((B, cap, 32), torch.uint8); b["cnt" + s] = pin((B,), torch.int32)
        b["u"] = pin((B, cap), torch.float32); b["z"] = pin((B, cap), torch.float32); b["n"] = pin((B,), torch.int32)
        b["bfi"] = pin((B, cap), torch.int32); b["bfd"] = pin((B, cap), torch.int32); b["bfs"] = pin((B, cap), torch.int32)
        hs.append((exL, exR, b))
    cap = hs[0][0].cap
    lib = m.lib()

    def enqueue(exL, exR, b):
        for ex, s in ((exL, "L"), (exR, "R")):
            m._check(lib.ygz_orb_extract_async(ex.h, b[s].data_ptr(), B, C5_W * C5_H, C5_W, b["kps" + s].data_ptr(), b["desc" + s].data_ptr(),
                                               b["cnt" + s].data_ptr(), cap), "ygz_orb_extract_async")
        m._check(lib.ygz_stereo_batch_async(exL.h, exR.h, B, C5_MB, C5_MBF, b["u"].data_ptr(), b["z"].data_ptr(), b["n"].data_ptr(),
                                            b["bfi"].data_ptr(), b["bfd"].data_ptr(), b["bfs"].data_ptr()), "ygz_stereo_batch_async")

    def wait(exL, exR):
        m._check(lib.ygz_orb_wait(exR.h), "ygz_orb_wait")
        m._check(lib.ygz_orb_wait(exL.h), "ygz_orb_wait")

    def e2e(steps):
        for i in range(steps):
            exL, exR, b = hs[i % npairs]
            if i >= npairs:
                wait(exL, exR)
            enqueue(exL, exR, b)
        for exL, exR, _ in hs:
            wait(exL, exR)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    e2e(max(args.warmup, npairs))
    exL, exR, b0 = hs[0]
    n_match = int(b0["n"].numpy().sum()); n_kp = int(b0["cntL"].numpy().sum() + b0["cntR"].numpy().sum())
    assert int(hs[1][2]["n"].numpy().sum()) == n_match
    for _ in range(args.warmup):
        m.time_stereo_frontend(exL, exR, B, 1, C5_MB, C5_MBF)
    sampler = ctx["ClockSampler"](local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    ms_dev, stage_last = m.time_stereo_frontend(exL, exR, B, args.steps, C5_MB, C5_MBF)
    barrier()
    t0 = time.perf_counter()
    e2e(args.steps)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier()
    clocks = sampler.finish()
    it = max(3, min(args.steps, 10))
    st = exL.time(B, it)
    stages = {"pyramid(7 launches)": st["pyramid"] / it, "fast_cells": st["fast_cells"] / it, "octree": st["octree"] / it,
              "blur_levels": st["blur"] / it, "orient_describe": st["describe"] / it}
    shard = ctx["shard"]
    ctr = exL.counters()
    ctr["matches"] = n_match
    ctr["elapsed_ns"] = int(ms_dev * 1e6)
    gathered = shard.gather_counters(dist, ctr, device="cuda" if dist is not None else None, timings_s=[ms_dev / 1e3, t_e2e])
    t_dev, t_e = shard.timings_max(gathered, 2)
    if rank != 0:
        return
    total = world * B * args.steps
    peak, peak_src = ctx["measured_peak_gbs"]()
    sizes = _level_sizes(C5_W, C5_H, 8, 1.2)
    sum_p = sum(w * h for w, h in sizes)
    fast_ms = stages["fast_cells"]
    launches_per_step = 2 * (st["launches"] // it) + 3
    out = {
        "metric": "C5 stereo front-end pairs/sec (1242x375 x2, 2000 feat, extractor x2 + ComputeStereoMatches + brute-force matcher)",
        "value": total / t_dev, "unit": "stereo pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": c5_config(B),
        "e2e": {"value": total / t_e, "unit": "stereo pairs/s", "h2d_bytes_per_step": int(2 * B * C5_W * C5_H),
                "d2h_bytes_per_step": int(2 * (B * cap * (28 + 32) + B * 4) + B * cap * 4 * 5 + B * 4), "ms_per_step": t_e / args.steps * 1e3,
                "pipeline": "%d batches in flight (ygz_orb_extract_async x2 + ygz_stereo_batch_async + ygz_orb_wait)" % npairs},
        "gpu_launches": int(args.steps * launches_per_step), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "fast_cells (one of the two extractors)", "achieved": sum_p * B / (fast_ms / 1e3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": sum_p * B / (fast_ms / 1e3) / 1e9 / peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": sum_p * B, "kernel_ms_per_launch": fast_ms,
                     "limiter": "issue-bound like the C2 kernel (profiles/kernel_metrics.json is captured on the C2 step)"},
        "stages_ms_per_step": {"extract_L_and_R_concurrent": stage_last["extract"], "stereo_match": stage_last["match"],
                               "median_cut": stage_last["median_cut"], "brute_force": stage_last["brute_force"],
                               "one_extractor": stages},
        "keypoints_per_image": n_kp / (2 * B), "stereo_matches_per_pair": n_match / B,
        "counters_all_ranks": {"columns": list(shard.COUNTER_COLUMNS), "rows": gathered[:, :len(shard.COUNTER_COLUMNS)].tolist()},
    }
    if not args.no_cpu_baseline and world == 1:
        threads, note = ctx["host_threads"]()
        ns = min(B, max(2 * threads, 16))
        res = m.StereoBatch(exL, exR, B, C5_MB, C5_MBF, brute_force=False)
        fps, done, dt, native, gate = c5_cpu_arm(Limg[:ns], Rimg[:ns], threads, 8.0, check=res)
        out["cpu_baseline"] = {"value": fps, "unit": "stereo pairs/s", "cores": threads, "kind": "port",
                               "sample": "%d pairs of the same set in %.1f s, oracle port (%s build), %d threads (%s)" %
                                         (done, dt, "-march=native" if native else "portable", threads, note),
                               "parity_gate_pair0_bit_exact": gate}
        f1, d1, t1, _, _ = c5_cpu_arm(Limg[:4], Rimg[:4], 1, 4.0)
        out["cpu_baseline"]["value_1_thread"] = f1
    ctx["emit_json"](out)


def run_c5_reference(args, B, ctx):
    threads, note = ctx["host_threads"]()
    Limg, Rimg = c5_pairs(1, B)
    t0 = time.perf_counter()
    fps, done, dt, native, _ = c5_cpu_arm(Limg, Rimg, threads, 0.0)
    out = {"impl": "reference", "metric": "C5 stereo front-end pairs/sec (1242x375 x2, 2000 feat, extractor x2 + ComputeStereoMatches + brute-force matcher)",
           "value": fps, "unit": "stereo pairs/s", "n_gpus": args.gpus, "steps": 1, "warmup": 1, "ms_per_step": dt * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": c5_config(B),
           "cpu_baseline": {"value": fps, "unit": "stereo pairs/s", "cores": threads, "kind": "port",
                            "sample": "%d pairs, oracle port (%s build), %d threads (%s)" % (done, "-march=native" if native else "portable", threads, note)},
           "e2e": {"value": fps, "unit": "stereo pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    ctx["emit_json"](out)


# ----------------------------------------------------------------------------------------------------------------- C3
def c3_config(P):
    return {"workload": "C3 SparseImgAlign::run: 752x480 ref/cur pair rendered from 200 points of known depth, 4-level x2.0 pyramid, levels "
                        "3..1, 4x4 patches, 10 Gauss-Newton iterations per level, eps 1e-6, EuRoC intrinsics, identity initial guess",
            "problems_per_step": P, "patches": C3_PATCHES,
            "l2_policy": "latency-bound iterative solver: %d problems x 237 KB of level images = %.0f MB touched per step" % (P, P * 0.237),
            "parallelism": "replicas: independent alignment problems per GPU"}


def c3_problems(m, n_distinct, device, cls=None):
    synth = _synth()
    px = m.ORBextractor(100, 2.0, C3_LEVELS, 20, 7, width=C3_W, height=C3_H, max_batch=1, device=device)
    inv = px.tables()["inv_scale"]
    probs, scenes = [], []
    for seed in range(n_distinct):
        s = synth.synth_sia_pair(seed, n_pts=C3_PATCHES, w=C3_W, h=C3_H)
        px.ComputePyramid(s["ref"]); ref = [px.level(0, l) for l in range(C3_LEVELS)]
        px.ComputePyramid(s["cur"]); cur = [px.level(0, l) for l in range(C3_LEVELS)]
        probs.append(m.make_sia_problem(ref, cur, inv, s["K"], s["px"], s["xyz"], cls=cls))
        scenes.append((ref, cur, inv, s))
    return probs, scenes


def c3_cpu_arm(scenes, m, threads, min_seconds, gpu_poses=None):
    from concurrent.futures import ThreadPoolExecutor
    O, native = _oracle()
    probs = [m.make_sia_problem(ref, cur, inv, s["K"], s["px"], s["xyz"], cls=O.SiaProblem) for ref, cur, inv, s in scenes]

    def one(i):
        return O.sia_run(probs[i % len(probs)], C3_MAX, C3_MIN, C3_ITERS, 1e-6)

    gate = None
    if gpu_poses is not None:
        worst = 0.0
        for i in range(len(probs)):
            n, q, t, _, _ = one(i)
            gq, gt, gn = gpu_poses[i]
            dq = min(np.abs(q - gq).max(), np.abs(q + gq).max())
            worst = max(worst, float(dq), float(np.abs(t - gt).max()))
            if n != gn:
                worst = float("inf")
        gate = worst
    n = max(len(probs), 4 * threads)
    with ThreadPoolExecutor(threads) as pool:
        list(pool.map(one, range(threads)))
        done, t0 = 0, time.perf_counter()
        while True:
            list(pool.map(one, range(n)))
            done += n
            dt = time.perf_counter() - t0
            if dt >= min_seconds:
                break
    return done / dt, done, dt, native, gate


def run_c3(args, P, ctx):
    m, torch, dist, rank, world, local = ctx["m"], ctx["torch"], ctx["dist"], ctx["rank"], ctx["world"], ctx["local"]
    nd = 8
    probs, scenes = c3_problems(m, nd, local)
    sia = m.SparseImgAlign(C3_MAX, C3_MIN, C3_ITERS, device=local, max_problems=P, max_features=256, max_level_pixels=2 * C3_W * C3_H)
    batch = [probs[i % nd] for i in range(P)]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        q, t, H, nt = sia.run_batch(batch)
    iters_per_problem = [len(sia.trace(i)) for i in range(nd)]
    gn_iters = sum(iters_per_problem[i % nd] for i in range(P))
    sampler = ctx["ClockSampler"](local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    ms_dev = sia.time(P, args.steps)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sia.run_batch(batch)                                  # host problem structs: level images H2D, poses D2H inside
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier()
    # tracker form: both frames were just extracted on this device, the aligner reads their pyramids in place (ygz_sia_run_orb)
    ex = m.ORBextractor(100, 2.0, C3_LEVELS, 20, 7, width=C3_W, height=C3_H, max_batch=2 * nd, device=local)
    ex.extract_batch([scenes[i // 2][3]["ref" if i % 2 == 0 else "cur"] for i in range(2 * nd)])
    feats = [m.make_sia_features(scenes[i][3]["K"], scenes[i][3]["px"], scenes[i][3]["xyz"]) for i in range(nd)]
    fbatch = [feats[i % nd] for i in range(P)]
    rf = np.array([2 * (i % nd) for i in range(P)], np.int32)
    cf = rf + 1
    qo, to, _, nto = sia.run_batch_orb(ex, ex, fbatch, rf, cf)
    orb_same = bool(np.array_equal(qo, q) and np.array_equal(to, t) and np.array_equal(nto, nt))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sia.run_batch_orb(ex, ex, fbatch, rf, cf)
    torch.cuda.synchronize()
    t_orb = time.perf_counter() - t0
    barrier()
    clocks = sampler.finish()
    one = m.SparseImgAlign(C3_MAX, C3_MIN, C3_ITERS, device=local, max_problems=1, max_features=256, max_level_pixels=2 * C3_W * C3_H)
    one.run(probs[0])
    lat_ms = one.time(1, 50) / 50
    shard = ctx["shard"]
    ctr = dict(frames=P * args.steps, keypoints=0, matches=0, sia_iters=gn_iters * args.steps, bytes_in=0, bytes_out=0,
               kernel_launches=args.steps, elapsed_ns=int(ms_dev * 1e6))
    gathered = shard.gather_counters(dist, ctr, device="cuda" if dist is not None else None, timings_s=[ms_dev / 1e3, t_e2e, t_orb])
    t_dev, t_e, t_o = shard.timings_max(gathered, 3)
    if rank != 0:
        return
    total = world * P * args.steps
    peak, peak_src = ctx["measured_peak_gbs"]()
    bytes_iter = C3_PATCHES * (25 + 64 + 384)                  # 94.6 KB per iteration at N = 200
    alg = gn_iters * bytes_iter
    img_bytes = sum(w * h for w, h in _level_sizes(C3_W, C3_H, C3_LEVELS, 2.0)[1:]) * 2
    out = {
        "metric": "C3 SparseImgAlign alignments/sec (4-level pyramid, 200 patches, levels 3..1, 10 GN iterations per level)",
        "value": total / t_dev, "unit": "alignments/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": c3_config(P),
        "gn_iterations_per_s": world * gn_iters * args.steps / t_dev, "gn_iterations_per_alignment": gn_iters / P,
        "single_problem_latency_ms": lat_ms,
        "e2e": {"value": total / t_e, "unit": "alignments/s", "h2d_bytes_per_step": int(P * img_bytes + P * C3_PATCHES * 21),
                "d2h_bytes_per_step": int(P * (7 + 36 + 1) * 4), "ms_per_step": t_e / args.steps * 1e3,
                "pipeline": "ygz_sia_run with host ygz_sia_problem structs (level images uploaded every call)"},
        "e2e_device_pyramids": {"value": total / t_o, "unit": "alignments/s", "h2d_bytes_per_step": int(P * C3_PATCHES * 21 + P * 264),
                                "d2h_bytes_per_step": int(P * (7 + 36 + 1) * 4), "ms_per_step": t_o / args.steps * 1e3,
                                "same_result_as_host_images": orb_same,
                                "pipeline": "ygz_sia_run_orb: level images read in place from the ORBextractor handle that extracted the two frames "
                                            "(their upload is the extractor's, measured in the C2 / C5 e2e); per step: features + poses H2D, "
                                            "one kernel, poses + H D2H"},
        "gpu_launches": int(args.steps), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_sia_align", "achieved": alg / (t_dev / args.steps) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (t_dev / args.steps) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg,
                     "limiter": "latency-bound iterative solver (SURVEY.md §8d: report alignments/s and iterations/s, no roofline target)"},
        "counters_all_ranks": {"columns": list(shard.COUNTER_COLUMNS), "rows": gathered[:, :len(shard.COUNTER_COLUMNS)].tolist()},
    }
    if not args.no_cpu_baseline and world == 1:
        threads, note = ctx["host_threads"]()
        gp = [(q[i], t[i], int(nt[i])) for i in range(nd)]
        fps, done, dt, native, gate = c3_cpu_arm(scenes, m, threads, 8.0, gpu_poses=gp)
        f1, d1, t1, _, _ = c3_cpu_arm(scenes, m, 1, 4.0)
        out["cpu_baseline"] = {"value": fps, "unit": "alignments/s", "cores": threads, "kind": "port", "value_1_thread": f1,
                               "sample": "%d alignments in %.1f s, oracle port (%s build), %d threads (%s); %d in %.1f s on 1 thread" %
                                         (done, dt, "-march=native" if native else "portable", threads, note, d1, t1),
                               "parity_gate_max_pose_diff": gate, "parity_gate_bound": 1e-4}
    ctx["emit_json"](out)


def run_c3_reference(args, P, ctx):
    import __graft_entry__ as g
    m = g.load_package()
    threads, note = ctx["host_threads"]()
    synth = _synth()
    O, native = _oracle()
    ox = O.Extractor(100, 2.0, C3_LEVELS, 20, 7)
    inv = ox.tables()["inv_scale"]
    scenes = []
    for seed in range(8):
        s = synth.synth_sia_pair(seed, n_pts=C3_PATCHES, w=C3_W, h=C3_H)
        ox.extract(s["ref"]); ref = [ox.level(l) for l in range(C3_LEVELS)]
        ox.extract(s["cur"]); cur = [ox.level(l) for l in range(C3_LEVELS)]
        scenes.append((ref, cur, inv, s))
    fps, done, dt, native, _ = c3_cpu_arm(scenes, m, threads, 6.0)
    out = {"impl": "reference", "metric": "C3 SparseImgAlign alignments/sec (4-level pyramid, 200 patches, levels 3..1, 10 GN iterations per level)",
           "value": fps, "unit": "alignments/s", "n_gpus": args.gpus, "steps": 1, "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": c3_config(P),
           "cpu_baseline": {"value": fps, "unit": "alignments/s", "cores": threads, "kind": "port",
                            "sample": "%d alignments in %.1f s, oracle port (%s build), %d threads (%s)" %
                                      (done, dt, "-march=native" if native else "portable", threads, note)},
           "e2e": {"value": fps, "unit": "alignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    ctx["emit_json"](out)
