"""One bench-shaped C2 step for ncu: extractor (B frames) + Frame::ComputeBoW through the ORBvoc-shaped vocabulary + batched
SearchByBoW.  Two warm-up steps, then exactly ONE step between cudaProfilerStart/Stop: run ncu with --profile-from-start off
and no launch-skip arithmetic is needed.  usage: python profiles/scripts/prof1.py [B]"""
import ctypes, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
import bench
m = g.load_package(); import importlib
synth = importlib.import_module('orb_ygz_slam_b200.synth')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
imgs = synth.synth_stream(1000, B)
ex = m.ORBextractor(1000, 1.2, 8, 20, 7, width=752, height=480, max_batch=B)
mt = m.ORBmatcher(0.7, True)
voc = m.ORBVocabulary(data=bench.vocabulary_bytes(), device=0)
mt.set_vocabulary(voc, bench.LEVELSUP)
ex.upload(imgs)
kf = np.array([(i - 1) % B for i in range(B)], np.int32); f = np.arange(B, dtype=np.int32)
rt = ctypes.CDLL('/usr/local/cuda/lib64/libcudart.so')
for it in range(3):
    if it == 2:
        rt.cudaProfilerStart()
    ex.run(B); ex.sync()
    nm, out = mt.SearchByBoWBatch(ex, kf, f, None)
    if it == 2:
        rt.cudaDeviceSynchronize(); rt.cudaProfilerStop()
print('done', nm.mean())
