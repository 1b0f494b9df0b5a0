import sys, time; sys.path[:0]=['.']
import numpy as np, torch
import bench
from fetalsyngen_b200.utils.phantom import label_phantom
shape=(256,256,256); DEV='cuda:0'
seg_h, seeds_h = label_phantom(shape)
gen = bench.build_generator(shape, DEV)
seg_d = torch.from_numpy(seg_h).to(DEV); seeds_d=[torch.from_numpy(s).to(DEV) for s in seeds_h]
img, seg, _ = gen.sample_batch([seg_d],[seeds_d],scale=False,sample_ids=[0],base_seed=1)
out, segf = img[0], seg[0].float()
art = bench.default_artifacts(1.0)["simulate_motion"]
ts=[]
for r in range(14):
    np.random.seed(r); torch.manual_seed(r)
    torch.cuda.synchronize(); t0=time.perf_counter()
    y, meta = art(out, segf, DEV, {}, resolution=[0.5]*3)
    t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
    ts.append(((t1-t0)*1e3,(t2-t0)*1e3, meta["nstacks"]))
for a in ts: print("host %.1f ms total %.1f ms nstacks %s" % a)
import cProfile, pstats, io
pr=cProfile.Profile(); pr.enable()
for r in range(4):
    np.random.seed(30+r); torch.manual_seed(30+r); art(out, segf, DEV, {}, resolution=[0.5]*3)
torch.cuda.synchronize(); pr.disable()
s=io.StringIO(); st=pstats.Stats(pr,stream=s).sort_stats('tottime'); st.print_stats(6); st.print_callers("method 'to' of"); st.print_callers("method 'cpu' of"); print(s.getvalue()[:6000])
