import sys, time; sys.path[:0]=['.']
import numpy as np, torch
import bench
from fetalsyngen_b200 import _lib
from fetalsyngen_b200.utils.phantom import label_phantom
shape=(256,256,256); DEV='cuda:0'
seg_h, seeds_h = label_phantom(shape)
gen = bench.build_generator(shape, DEV)
seg_d = torch.from_numpy(seg_h).to(DEV); seeds_d=[torch.from_numpy(s).to(DEV) for s in seeds_h]
img, seg, _ = gen.sample_batch([seg_d],[seeds_d],scale=False,sample_ids=[0],base_seed=1)
out, segf = img[0], seg[0].float()
arts = bench.default_artifacts(1.0)
for name in ("boundaries","blur_cortex","struct_noise"):
    art = arts[name]
    for r in range(3):
        np.random.seed(r); torch.manual_seed(r); art(out, segf, DEV, {}, resolution=[0.5]*3)
    torch.cuda.synchronize()
    for r in range(3):
        np.random.seed(10+r); torch.manual_seed(10+r)
        _lib.stats.reset(); _lib.stats.timing=True
        t0=time.perf_counter(); y, meta = art(out, segf, DEV, {}, resolution=[0.5]*3); t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
        _lib.stats.timing=False
        per=_lib.stats.elapsed_ms()
        print(name, "host ms %.2f wall ms %.2f" % ((t1-t0)*1e3,(t2-t0)*1e3), "kernels ms %.2f" % sum(v[1] for v in per.values()), {k:(v[0], round(v[1],2)) for k,v in sorted(per.items(), key=lambda kv:-kv[1][1])[:7]})
