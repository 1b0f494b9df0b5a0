import os, sys; sys.path[:0]=['.']
os.environ["CUDA_LAUNCH_BLOCKING"]="1"
import numpy as np, torch
import bench
from fetalsyngen_b200.utils.phantom import label_phantom
from fetalsyngen_b200.sharding import step_ids
S=int(sys.argv[1]); B=int(sys.argv[2]); steps=int(sys.argv[3]) if len(sys.argv)>3 else 2
shape=(S,S,S); dev='cuda:0'
seg_h, seeds_h = label_phantom(shape)
gen = bench.build_generator(shape, dev)
seg_d = torch.from_numpy(seg_h).to(dev); seeds_d=[torch.from_numpy(s).to(dev) for s in seeds_h]
for st in range(steps):
    ids=step_ids(st,B,0,1)
    try:
        img, seg, params = gen.sample_batch([seg_d]*B,[seeds_d]*B,scale=True,sample_ids=ids,base_seed=1234)
        torch.cuda.synchronize()
        print("step", st, "ok", float(img.mean()), [np.round(np.asarray(p['deform_params']['affine']['rotations']),2).tolist() for p in params][:2])
    except Exception as e:
        print("step", st, "FAILED", str(e)[:300]); break
