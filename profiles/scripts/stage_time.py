import sys; sys.path.insert(0,".")
import __graft_entry__ as g, numpy as np, importlib
m=g.load_package(); synth=importlib.import_module("orb_ygz_slam_b200.synth")
B=256; ex=m.ORBextractor(1000,1.2,8,20,7,width=752,height=480,max_batch=B); imgs=synth.synth_stream(1000,B); ex.upload(imgs); ex.run(B); ex.sync()
t=ex.time(B,10); print({k:(round(v/10,4) if k!="launches" else v) for k,v in t.items()})
