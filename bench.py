#!/usr/bin/env python
"""bench.py — front-end frames/sec of the B200-native ORB-YGZ-SLAM image front end.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--batch B]

Workload (BASELINE.json configs[1], "C2"): synthetic EuRoC-shape mono stream, 752x480 u8 frames,
ORBextractor(1000, 1.2, 8, 20, 7) + ORBmatcher(0.7, checkOri).SearchByBoW of every frame against the previous
one (synthetic bench vocabulary node = descriptor byte0 >> 2; DBoW2 transform is a "next" row).
One "step" = one batch of B frames through the whole hot path (7 pyramid launches, FAST cells, octree,
orientation+rBRIEF, batched SearchByBoW).  Inputs are B distinct frames (B*361 KB = 92 MB at B=256; pyramid working
set 286 MB > the 126 MB L2, so no L2 flush is needed between steps).

  value  whole-job frames/s with inputs already resident in HBM (CUDA events on the launching stream)
  e2e    the same through the public C-ABI calls with HOST (pinned) buffers: H2D of the frames, D2H of
         keypoints + descriptors + match arrays inside the timed region
  roofline  for the dominant kernel (by device time) against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline  the oracle port (host-tuned build, all host threads) on a bounded sample of the same workload

Multi-GPU (torchrun, one rank per GPU): frames are independent units, every rank runs its own batch (weak
scaling, no data-path collective); the only exchange is one NCCL all-gather of throughput counters.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 752, 480
ORB = dict(nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7)
NNRATIO, CHECK_ORI = 0.7, True
# algorithmic bytes per frame (SURVEY.md §8d, C1 geometry)
LEVELS = [(752, 480), (627, 400), (522, 333), (435, 278), (363, 231), (302, 193), (252, 161), (210, 134)]
SUM_P = sum(w * h for w, h in LEVELS)                      # 1 117 367 : read every level once (FAST) / strict pyramid+FAST
PYR_CHAIN = sum(w * h for w, h in LEVELS[:-1]) + sum(w * h for w, h in LEVELS[1:])   # 1 845 634 : per-level resize chain


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                p = [x.strip() for x in line.split(",")]
                if len(p) >= 7:
                    self.samples.append(p)
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for p in self.samples:
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Host threads the CPU arm can actually use: logical CPUs, limited by affinity and by the cgroup CPU quota
    (the GPU boxes expose 128 logical CPUs under a 16-CPU quota; oversubscribing a quota only adds throttling)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    note = "%d logical CPUs" % (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = max(1, int(round(int(q) / int(per))))
            if quota < n:
                n = quota
                note += ", cgroup cpu.max quota %d" % quota
    except Exception:
        pass
    return n, note


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_frames(seed, n):
    import __graft_entry__ as g
    g.load_package()
    import importlib
    synth = importlib.import_module("orb_ygz_slam_b200.synth")
    return synth.synth_stream(seed, n, W, H)


def cpu_arm(frames, threads, min_seconds=8.0, max_frames=None):
    """Oracle port (host-tuned build) on all host threads over a bounded sample; returns (fps, n_frames, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    try:
        O.lib(native=True)
        native = True
    except Exception:
        native = False
    n = len(frames) if max_frames is None else min(len(frames), max_frames)
    O.frontend_batch(frames[:min(n, max(threads, 8))], threads, nnratio=NNRATIO, check_ori=CHECK_ORI, native=native, **_orb_kw())  # warm
    # every thread keeps its chunk of the batch and re-processes it `repeat` times inside ONE call, so thread start-up and
    # first-touch page faults are amortised (a per-call spawn of 128 threads over 2 frames each measured 10x too slow)
    repeat = 4
    done, t0 = 0, time.perf_counter()
    while True:
        O.frontend_batch(frames[:n], threads, nnratio=NNRATIO, check_ori=CHECK_ORI, native=native, repeat=repeat, **_orb_kw())
        done += n * repeat
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
        repeat = min(repeat * 2, 64)
    return done / dt, done, dt, native


def _orb_kw():
    return dict(nfeatures=ORB["nfeatures"], scale=ORB["scaleFactor"], nlevels=ORB["nlevels"], ini_th=ORB["iniThFAST"],
                min_th=ORB["minThFAST"])


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference itself cannot be
    compiled here — no OpenCV/Eigen headers) on all host threads.  Rank 0 only."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    threads, cpu_note = host_threads()
    sample = max(threads, 64)
    frames = make_frames(1234, min(sample, 256))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    try:
        O.lib(native=True); native = True
    except Exception:
        native = False
    kw = dict(nnratio=NNRATIO, check_ori=CHECK_ORI, native=native, **_orb_kw())
    O.frontend_batch(frames, threads, repeat=max(args.warmup, 1), **kw)
    t0 = time.perf_counter()
    O.frontend_batch(frames, threads, repeat=args.steps, **kw)      # K steps inside one call: worker threads persist across steps
    dt = time.perf_counter() - t0
    fps = args.steps * len(frames) / dt
    out = {
        "impl": "reference", "metric": "front-end frames/sec (752x480, 1000 feat, 8-lvl pyr)", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(len(frames)),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d frames/step x %d steps, oracle port (%s build), %d threads (%s)" %
                                   (len(frames), args.steps, "-march=native" if native else "portable", threads, cpu_note)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(out)
    return 0


def aux_metrics(m, device):
    """Secondary configurations of BASELINE.json (not part of `value`): C3 SparseImgAlign batch and C5 brute-force
    stereo-size matcher, device-timed with CUDA events."""
    import importlib
    synth = importlib.import_module("orb_ygz_slam_b200.synth")
    aux = {}
    try:
        # C3: 4-level x2.0 pyramid (built by the device extractor), 200 patches, levels 3..1, 10 GN iterations
        px = m.ORBextractor(1000, 2.0, 4, 20, 7, width=W, height=H, max_batch=1, device=device)
        inv = px.tables()["inv_scale"]
        probs = []
        for seed in range(4):
            s = synth.synth_sia_pair(seed, n_pts=200, w=W, h=H)
            px.ComputePyramid(s["ref"]); ref = [px.level(0, l) for l in range(4)]
            px.ComputePyramid(s["cur"]); cur = [px.level(0, l) for l in range(4)]
            probs.append(m.make_sia_problem(ref, cur, inv, s["K"], s["px"], s["xyz"]))
        P = 64
        sia = m.SparseImgAlign(3, 1, 10, device=device, max_problems=P, max_features=256, max_level_pixels=2 * W * H)
        batch = [probs[i % 4] for i in range(P)]
        sia.run_batch(batch)
        it = 20
        ms = sia.time(P, it) / it
        aux["sia_c3"] = {"problems_per_launch": P, "patches": 200, "levels": "3..1", "gn_iters_per_level": 10,
                         "ms_per_launch": ms, "alignments_per_s": P / (ms / 1e3)}
    except Exception as e:        # aux numbers never break the headline line
        aux["sia_c3"] = {"error": str(e)[:200]}
    try:
        q, t, _ = synth.synth_descriptors(5, 2000, 2000)
        mt = m.ORBmatcher(0.7, True, device=device)
        import time as _t
        mt.bf_top2(q, t)
        t0 = _t.perf_counter()
        for _ in range(20):
            mt.bf_top2(q, t)
        dt = (_t.perf_counter() - t0) / 20
        aux["bf_top2_c5"] = {"nq": 2000, "nt": 2000, "ms_per_call_host_buffers": dt * 1e3, "comparisons_per_s": 4e6 / dt}
    except Exception as e:
        aux["bf_top2_c5"] = {"error": str(e)[:200]}
    try:
        # FindDirectProjection over the (map point, keyframe) items of a batch of frames
        import time as _t
        s = synth.synth_direct_scene(9, n_kf=3, n_items=16384, w=W, h=H)
        px = m.ORBextractor(1000, 1.2, 8, 20, 7, width=W, height=H, max_batch=1, device=device)
        tb = px.tables()

        def pyr(img):
            px.ComputePyramid(img)
            return [px.level(0, l) for l in range(8)]
        frames = [m.make_direct_frame(pyr(im), q, t) for im, (q, t) in zip(s["kf_imgs"], s["kf_poses"])]
        frames.append(m.make_direct_frame(pyr(s["cur_img"]), *s["cur_pose"]))
        dm = m.DirectMatcher(s["K"], tb["scale"], tb["inv_scale"], tb["inv_sigma2"][1], max_slots=4, max_items=16384,
                             max_pyramid_bytes=4 * W * H, device=device)
        for k, f in enumerate(frames):
            dm.upload_frame(k, f)
        dm.FindDirectProjection(3, s["items"])
        t0 = _t.perf_counter()
        for _ in range(10):
            _, _, ok = dm.FindDirectProjection(3, s["items"])
        dt = (_t.perf_counter() - t0) / 10
        ms = dm.time(20) / 20
        aux["direct_projection"] = {"items_per_call": 16384, "ms_kernel": ms, "ms_per_call_host_buffers": dt * 1e3,
                                    "items_per_s": 16384 / dt, "success_rate": float(ok.mean())}
    except Exception as e:
        aux["direct_projection"] = {"error": str(e)[:200]}
    try:
        # Frame::ComputeBoW for a batch: vocabulary descent of every descriptor the extractor holds on the device
        voc = m.ORBVocabulary(data=synth.synth_vocabulary(1, k=10, L=5), device=device)
        Bv = 64
        exv = m.ORBextractor(width=W, height=H, max_batch=Bv, device=device, **ORB)
        exv.upload(make_frames(77, Bv)); exv.run(Bv); exv.sync()
        voc.transform_orb(exv, Bv, 4, want=(), timed=True)
        r = voc.transform_orb(exv, Bv, 4, want=(), timed=True)
        nd = int(exv.download(Bv)[2].sum())
        aux["vocab_transform"] = {"vocabulary": "synthetic k=10 L=5 (101 k nodes)", "frames": Bv, "descriptors": nd,
                                  "ms_kernel": r["ms"], "descriptors_per_s": nd / (r["ms"] / 1e3)}
    except Exception as e:
        aux["vocab_transform"] = {"error": str(e)[:200]}
    return aux


def workload_config(batch):
    return {"workload": "C2 synthetic EuRoC-shape mono stream 752x480 u8: ORBextractor(1000,1.2,8,20,7) + "
                        "ORBmatcher(0.7,checkOri).SearchByBoW vs previous frame (synthetic vocabulary byte0>>2)",
            "frames_per_step": batch, "l2_policy": "inputs larger than L2 (%d distinct frames/step, pyramid working set %.0f MB)"
                                                   % (batch, batch * (SUM_P + 361e3 * 0) / 1e6),
            "in_flight": "2 independent batches per GPU on 2 streams for the device-resident `value`; `e2e.pipeline` states its own depth",
            "parallelism": "replicas: one independent batch per GPU"}


_JSON_FD = None


def emit_json(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, line)
    else:
        os.write(_JSON_FD, line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true")
    ap.add_argument("--inflight", type=int, default=3, help="batches in flight on the end-to-end (host buffer) path")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # The contract is ONE JSON line on stdout.  Native libraries (NCCL prints its version banner there) share fd 1, so
    # everything except the final line goes to stderr: fd 1 is pointed at fd 2 for the whole run and the JSON line is
    # written to the saved descriptor.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as g
    g.build_cuda()
    m = g.load_package()
    if m.device_count() <= local:
        raise SystemExit("bench.py needs a CUDA device (product path has no CPU fallback)")
    B = args.batch
    frames = make_frames(1000 + rank, B)

    ex = m.ORBextractor(width=W, height=H, max_batch=B, device=local, **ORB)
    mt = m.ORBmatcher(NNRATIO, CHECK_ORI, device=local)
    # second handle pair: the end-to-end path alternates two in-flight batches so one batch's H2D/D2H overlaps the other's kernels
    ex2 = m.ORBextractor(width=W, height=H, max_batch=B, device=local, **ORB)
    mt2 = m.ORBmatcher(NNRATIO, CHECK_ORI, device=local)
    cap = ex.cap
    kf_idx = np.array([(i - 1) % B for i in range(B)], np.int32)
    f_idx = np.arange(B, dtype=np.int32)

    # pinned host buffers (torch is plumbing: pinned memory + the one collective)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
    h_img = pin((B, H, W), torch.uint8)
    h_img.numpy()[:] = frames
    h_kps = pin((B, cap, 7), torch.float32)
    h_desc = pin((B, cap, 32), torch.uint8)
    h_cnt = pin((B,), torch.int32)
    h_match = pin((B, cap), torch.int32)
    h_nm = pin((B,), torch.int32)
    bufs2 = dict(img=pin((B, H, W), torch.uint8), kps=pin((B, cap, 7), torch.float32), desc=pin((B, cap, 32), torch.uint8),
                 cnt=pin((B,), torch.int32), match=pin((B, cap), torch.int32), nm=pin((B,), torch.int32))
    bufs2["img"].numpy()[:] = frames
    bufs1 = dict(img=h_img, kps=h_kps, desc=h_desc, cnt=h_cnt, match=h_match, nm=h_nm)
    pairs = [(ex, mt, bufs1), (ex2, mt2, bufs2)]
    for _ in range(max(0, args.inflight - 2)):          # further handle pairs for the end-to-end pipeline
        exn = m.ORBextractor(width=W, height=H, max_batch=B, device=local, **ORB)
        mtn = m.ORBmatcher(NNRATIO, CHECK_ORI, device=local)
        bn = dict(img=pin((B, H, W), torch.uint8), kps=pin((B, cap, 7), torch.float32), desc=pin((B, cap, 32), torch.uint8),
                  cnt=pin((B,), torch.int32), match=pin((B, cap), torch.int32), nm=pin((B,), torch.int32))
        bn["img"].numpy()[:] = frames
        pairs.append((exn, mtn, bn))
    pairs = pairs[:max(1, args.inflight)]
    import ctypes as C

    def enqueue(exh, mth, b):
        m._check(m.lib().ygz_orb_extract_async(exh.h, b["img"].data_ptr(), B, W * H, W, b["kps"].data_ptr(), b["desc"].data_ptr(),
                                               b["cnt"].data_ptr(), cap), "ygz_orb_extract_async")
        m._check(m.lib().ygz_search_by_bow_batch_async(mth.h, exh.h, NNRATIO, int(CHECK_ORI), kf_idx.ctypes.data_as(m.i32p),
                                                       f_idx.ctypes.data_as(m.i32p), B, b["match"].data_ptr(), b["nm"].data_ptr()),
                 "ygz_search_by_bow_batch_async")

    def e2e_pipelined(steps):
        np_ = len(pairs)
        for i in range(steps):
            exh, mth, b = pairs[i % np_]
            if i >= np_:
                m._check(m.lib().ygz_orb_wait(exh.h), "ygz_orb_wait")      # results of step i - inflight are on the host now
            enqueue(exh, mth, b)
        for exh, _, _ in pairs:
            m._check(m.lib().ygz_orb_wait(exh.h), "ygz_orb_wait")

    def e2e_step():
        m._check(m.lib().ygz_orb_extract(ex.h, h_img.data_ptr(), B, W * H, W, h_kps.data_ptr(), h_desc.data_ptr(),
                                         h_cnt.data_ptr(), cap), "ygz_orb_extract")
        m._check(m.lib().ygz_search_by_bow_batch(mt.h, ex.h, NNRATIO, int(CHECK_ORI), kf_idx.ctypes.data_as(m.i32p),
                                                 f_idx.ctypes.data_as(m.i32p), B, None, C.cast(h_match.data_ptr(), m.i32p),
                                                 C.cast(h_nm.data_ptr(), m.i32p)), "ygz_search_by_bow_batch")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also makes inputs resident and sets up the pair lists)
    for _ in range(args.warmup):
        e2e_step()
    n_kp = int(h_cnt.numpy().sum())
    n_match = int(h_nm.numpy().sum())
    m._check(m.lib().ygz_orb_upload(ex2.h, h_img.data_ptr(), B, W * H, W), "ygz_orb_upload")
    ex2.run(B)
    mt2.SearchByBoWBatch(ex2, kf_idx, f_idx, None)          # prepares the resident pair lists of the second handle pair
    for exn, mtn, bn in pairs[2:]:
        m._check(m.lib().ygz_orb_upload(exn.h, h_img.data_ptr(), B, W * H, W), "ygz_orb_upload")
        exn.run(B)
        mtn.SearchByBoWBatch(exn, kf_idx, f_idx, None)
    for _ in range(args.warmup):
        mt.time_frontend(ex, B, 1)
        mt.time_frontend2(ex, mt2, ex2, B, 2)

    # ---- device-resident timing: K steps, CUDA events on the launching stream
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    # two independent batches in flight on two streams (handle pairs ex/mt and ex2/mt2, inputs resident): exactly K steps,
    # alternating A, B, A, ...; CUDA events bracket the region on the launching streams (ygz_frontend_time2)
    ms_single = mt.time_frontend(ex, B, args.steps)
    barrier()
    ms_dev = mt.time_frontend2(ex, mt2, ex2, B, args.steps)
    barrier()
    # ---- end-to-end timing through the public C-ABI calls with host buffers (copies inside)
    barrier()
    e2e_pipelined(2)                     # warm the second handle pair
    barrier()
    t0 = time.perf_counter()
    e2e_pipelined(args.steps)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    for _, _, bn in pairs[1:]:
        assert int(bn["cnt"].numpy().sum()) == n_kp and int(bn["nm"].numpy().sum()) == n_match   # pipelined results == serial results
    barrier()
    clocks = sampler.finish()

    # per-stage device times (same launches, separately timed) for the roofline of the dominant kernel
    it = max(3, min(args.steps, 10))
    st = ex.time(B, it)
    bow_ms = mt.time_bow_batch(ex, B, it) / it
    stages = {"pyramid(7 launches)": st["pyramid"] / it, "fast_cells": st["fast_cells"] / it, "octree": st["octree"] / it,
              "blur_levels": st["blur"] / it, "orient_describe": st["describe"] / it, "bow_match": bow_ms}

    # ---- max over ranks; the ONE collective: all-gather of throughput counters (orb-ygz-slam_b200/shard.py)
    import importlib
    shard = importlib.import_module("orb_ygz_slam_b200.shard")
    ctr = ex.counters()
    ctr["matches"] = n_match
    t_dev, t_e = shard.max_over_ranks(dist, [ms_dev / 1e3, t_e2e], device="cuda" if dist is not None else None)
    ctr["elapsed_ns"] = int(t_dev * 1e9)
    gathered = shard.gather_counters(dist, ctr, device="cuda" if dist is not None else None)

    if rank == 0:
        total_frames = world * B * args.steps
        value = total_frames / t_dev
        e2e_v = total_frames / t_e
        peak, peak_src = measured_peak_gbs()
        dom = max(stages, key=stages.get)
        alg = {"pyramid(7 launches)": PYR_CHAIN, "fast_cells": SUM_P, "octree": 0, "blur_levels": 2 * SUM_P,
               "orient_describe": 0, "bow_match": 0}
        # algorithmic bytes of gather-type kernels: keypoints * (31x31 raw patch + 512 blurred samples + 60 B out) ; BoW: descriptors of both frames
        kp_per_frame = n_kp / B
        alg["orient_describe"] = int(kp_per_frame * (31 * 31 + 512 + 28 + 32))
        alg["octree"] = int(4 * 16139 / 16)     # ~4 B per FAST candidate (read once); informative only
        alg["bow_match"] = int(2 * kp_per_frame * (32 + 28) + 4 * cap)
        dom_bytes = alg[dom] * B
        achieved = dom_bytes / (stages[dom] / 1e3) / 1e9
        pf_bytes = (SUM_P) * B
        pf_ms = stages["pyramid(7 launches)"] + stages["fast_cells"]
        # DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch) from the ncu --set full capture of this very
        # command shape (B=256), profiles/r1_full_B256_raw.csv; None for other shapes
        NCU_TRAFFIC_B256 = {"fast_cells": 271616768 + 31951360, "orient_describe": 520987904 + 19014144,
                            "octree": 70572544 + 53032448, "blur_levels": 290054656 + 252909568}
        traffic = NCU_TRAFFIC_B256.get(dom) if B == 256 else None
        out = {
            "metric": "front-end frames/sec (752x480, 1000 feat, 8-lvl pyr)", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(B),
            "value_single_stream": world * B * args.steps / (ms_single / 1e3),
            "e2e": {"value": e2e_v, "unit": "frames/s", "h2d_bytes_per_step": int(h_img.numel()),
                    "d2h_bytes_per_step": int(h_kps.numel() * 4 + h_desc.numel() + h_cnt.numel() * 4 + h_match.numel() * 4 + h_nm.numel() * 4),
                    "ms_per_step": t_e / args.steps * 1e3,
                    "pipeline": "%d batches in flight (one handle pair each, ygz_orb_extract_async + ygz_search_by_bow_batch_async + ygz_orb_wait)" % len(pairs)},
            "gpu_launches": int(args.steps * ((ORB["nlevels"] - 1) + 4 + 1)),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                         "limiter": "issue slots, not bytes: ncu smsp__issue_active 80 % (fast_cells; ALU pipe 61 %), 87 % (blur_levels, which streams 2 x SUM_P at 2.6 TB/s); see DESIGN.md §4",
                         "kernel_ms_per_launch": stages[dom],
                         "other_kernels": {k: {"algorithmic_bytes_per_launch": alg[k] * B, "ms_per_launch": stages[k],
                                               "achieved": alg[k] * B / (stages[k] / 1e3) / 1e9,
                                               "frac": alg[k] * B / (stages[k] / 1e3) / 1e9 / peak,
                                               "traffic": NCU_TRAFFIC_B256.get(k) if B == 256 else None}
                                           for k in ("blur_levels", "pyramid(7 launches)", "orient_describe") if stages.get(k)},
                         "pyramid_plus_fast": {"strict_bytes_per_step": pf_bytes, "ms_per_step": pf_ms,
                                               "achieved": pf_bytes / (pf_ms / 1e3) / 1e9,
                                               "frac": pf_bytes / (pf_ms / 1e3) / 1e9 / peak}},
            "stages_ms_per_step": stages,
            "keypoints_per_frame": kp_per_frame, "matches_per_frame": n_match / B,
            "counters_all_ranks": {"columns": ["frames", "keypoints", "matches", "sia_iters", "bytes_in", "bytes_out",
                                               "kernel_launches", "elapsed_ns"], "rows": gathered.tolist()},
        }
        if world == 1 and not args.no_aux:
            out["aux"] = aux_metrics(m, local)
        if not args.no_cpu_baseline and world == 1:
            threads, cpu_note = host_threads()
            fps, done, dt, native = cpu_arm(frames, threads, min_seconds=8.0)
            out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                                   "sample": "%d frames of the same stream in %.1f s, oracle port (%s build), %d threads (%s)" %
                                             (done, dt, "-march=native" if native else "portable", threads, cpu_note)}
        emit_json(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
