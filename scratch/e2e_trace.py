import sys, time; sys.path.insert(0,'.')
import numpy as np, torch
import bench
from fetalsyngen_b200.utils.phantom import label_phantom
from fetalsyngen_b200.host_pipeline import HostPipeline
shape=(256,256,256); dev='cuda:0'; B=8
seg_h, seeds_h = label_phantom(shape)
gen = bench.build_generator(shape, dev)
hp = HostPipeline(gen, B, depth=3)
hp.set_inputs([seg_h]*B, [seeds_h]*B)
hp.run(3)
torch.cuda.synchronize()
# monkeypatch submit to record events
ev=[]
orig=hp.submit
def submit(scale=True, **kw):
    t0=time.perf_counter()
    a=torch.cuda.Event(enable_timing=True); a.record(hp.s_in)
    orig(scale, **kw)
    b=torch.cuda.Event(enable_timing=True); b.record(hp.s_in)
    c=torch.cuda.Event(enable_timing=True); c.record(hp.s_out)
    k=torch.cuda.Event(enable_timing=True); k.record(torch.cuda.current_stream())
    ev.append((a,b,c,k,t0,time.perf_counter()))
hp.submit=submit
base=torch.cuda.Event(enable_timing=True); base.record(); 
t0=time.perf_counter()
hp.run(8)
torch.cuda.synchronize(); print("total ms", (time.perf_counter()-t0)*1e3)
for i,(a,b,c,k,h0,h1) in enumerate(ev):
    print(i, "host submit %.1f-%.1f ms" % ((h0-t0)*1e3,(h1-t0)*1e3), "H2D %.1f->%.1f" % (base.elapsed_time(a), base.elapsed_time(b)), "compute done %.1f" % base.elapsed_time(k), "D2H done %.1f" % base.elapsed_time(c))
