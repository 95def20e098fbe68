"""One C3-shaped SparseImgAlign launch for ncu (P problems, a CTA each).  Warm-up call, then ONE call between cudaProfilerStart/Stop.  usage: python profiles/scripts/prof_sia.py [P]"""
import ctypes, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
import bench_c3c5 as b3
m = g.load_package()
P = int(sys.argv[1]) if len(sys.argv) > 1 else 592
probs, scenes = b3.c3_problems(m, 8, 0)
sia = m.SparseImgAlign(b3.C3_MAX, b3.C3_MIN, b3.C3_ITERS, device=0, max_problems=P, max_features=256, max_level_pixels=2 * b3.C3_W * b3.C3_H)
batch = [probs[i % 8] for i in range(P)]
rt = ctypes.CDLL('/usr/local/cuda/lib64/libcudart.so')
sia.run_batch(batch)
rt.cudaProfilerStart()
sia.time(P, 1)
rt.cudaDeviceSynchronize(); rt.cudaProfilerStop()
print('done')
