import cProfile, pstats, sys, io, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from fetalsyngen_b200.utils.phantom import label_phantom
shape=(256,256,256); dev='cuda:0'; B=8
seg_h, seeds_h = label_phantom(shape)
gen = bench.build_generator(shape, dev)
seg_d = torch.from_numpy(seg_h).to(dev); seeds_d=[torch.from_numpy(s).to(dev) for s in seeds_h]
out_img = torch.empty((B,*shape),dtype=torch.float32,device=dev); out_seg=torch.empty((B,*shape),dtype=torch.uint8,device=dev)
from fetalsyngen_b200.sharding import step_ids
cnt=[0]
def step():
    ids=step_ids(cnt[0],B,0,1); cnt[0]+=1
    gen.sample_batch([seg_d]*B,[seeds_d]*B,scale=True,out_img=out_img,out_seg=out_seg,sample_ids=ids,base_seed=1234)
for _ in range(3): step()
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(10): step()
t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
print("host ms/step", (t1-t0)*100, "incl sync", (t2-t0)*100)
pr=cProfile.Profile(); pr.enable()
for _ in range(10): step()
pr.disable(); torch.cuda.synchronize()
s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats("tottime").print_stats(45); print(s.getvalue()[:9000])
