"""Where the time of ygz_orb_extract_frame_dso_batch goes (host buffers), against the per-frame entry.  usage: python profiles/scripts/dso_timing.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as g

m = g.load_package()
import importlib
synth = importlib.import_module("orb_ygz_slam_b200.synth")
W, H, B = 752, 480, 64
fr = np.stack([synth.synth_frame(100 + i, W, H) for i in range(B)])
ex = m.ORBextractor(1000, 1.2, 8, 20, 7, width=W, height=H, max_batch=B)
L = m.lib()
i32p = C.POINTER(C.c_int)
for cap in (21828, 4096):
    ne = np.zeros(B, np.int32)
    kps = np.zeros((B, cap), m.KP_DTYPE); desc = np.zeros((B, cap, 32), np.uint8); counts = np.zeros(B, np.int32)
    gs = np.full(B, -1, np.int32)
    for rep in range(4):
        t0 = time.perf_counter()
        rc = L.ygz_orb_extract_frame_dso_batch(ex.h, fr.ctypes.data, B, fr.strides[0], fr.strides[1], ne.ctypes.data_as(i32p), kps.ctypes.data,
                                               desc.ctypes.data, counts.ctypes.data_as(i32p), cap, gs.ctypes.data_as(i32p))
        dt = time.perf_counter() - t0
        print("batch cap=%d rep %d: rc=%d %.2f ms (%.0f frames/s), grid sizes %s, counts %d..%d" % (cap, rep, rc, dt * 1e3, B / dt, sorted(set(gs.tolist())), counts.min(), counts.max()))
one = m.ORBextractor(1000, 1.2, 8, 20, 7, width=W, height=H)
cap = 4096
kps = np.zeros(cap, m.KP_DTYPE); desc = np.zeros((cap, 32), np.uint8); cnt = C.c_int(); gg = C.c_int(-1)
for rep in range(3):
    t0 = time.perf_counter()
    for f in range(B):
        L.ygz_orb_extract_frame_dso(one.h, fr[f].ctypes.data, fr.strides[1], 0, kps.ctypes.data, desc.ctypes.data, C.byref(cnt), cap, C.byref(gg))
    dt = time.perf_counter() - t0
    print("per-frame rep %d: %.2f ms for %d frames (%.0f frames/s), grid %d" % (rep, dt * 1e3, B, B / dt, gg.value))
t0 = time.perf_counter(); ex.upload(fr); ex.sync(); print("upload+sync %d frames: %.2f ms" % (B, (time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter(); ex.upload(fr); ex.sync(); print("upload+sync %d frames: %.2f ms" % (B, (time.perf_counter() - t0) * 1e3))
