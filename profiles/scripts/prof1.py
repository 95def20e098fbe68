import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import __graft_entry__ as g
m=g.load_package(); import importlib
synth=importlib.import_module('orb_ygz_slam_b200.synth')
B=int(sys.argv[1]) if len(sys.argv)>1 else 128
imgs=synth.synth_stream(1000,B)
ex=m.ORBextractor(1000,1.2,8,20,7,width=752,height=480,max_batch=B)
mt=m.ORBmatcher(0.7,True)
ex.upload(imgs)
for _ in range(3):
    ex.run(B); ex.sync()
kf=np.array([(i-1)%B for i in range(B)],np.int32); f=np.arange(B,dtype=np.int32)
nm,out=mt.SearchByBoWBatch(ex,kf,f,None)
print('done',nm.mean())
