import sys, time; sys.path.insert(0,'.')
import numpy as np, torch
import bench
from fetalsyngen_b200.engine import SynthEngine
from fetalsyngen_b200.utils.phantom import label_phantom
from fetalsyngen_b200.sharding import step_ids
shape=(256,256,256); dev='cuda:0'; B=8
seg_h, seeds_h = label_phantom(shape)
gen = bench.build_generator(shape, dev)
seg_d = torch.from_numpy(seg_h).to(dev); seeds_d=[torch.from_numpy(s).to(dev) for s in seeds_h]
out_img = torch.empty((B,*shape),dtype=torch.float32,device=dev); out_seg=torch.empty((B,*shape),dtype=torch.uint8,device=dev)
engA = gen.engine(shape)
engB = SynthEngine(shape, gen.resolution, dev); engB.tables = engA.tables
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
cnt=[0]
def plans_for(ids):
    # reuse sample_batch's drawing by calling its internals
    from fetalsyngen_b200.engine import SamplePlan
    from fetalsyngen_b200.sharding import sample_seed
    plans=[]
    for i in ids:
        sd=sample_seed(1234,i); np.random.seed(sd); torch.default_generator.manual_seed(sd)
        p=SamplePlan(rng_seed=1234, sample_id=i)
        gen._draw_generate(p, None, shape, {}, None, device_grids=True)
        p.mus,p.sigmas=gen.intensity_generator.draw_gmm({})
        gen._draw_augment(p, shape, {}, None, device_grids=True)
        plans.append(p)
    return plans
def step_one():
    ids=step_ids(cnt[0],B,0,1); cnt[0]+=1
    pl=plans_for(ids)
    engA.run_base(pl,[[s.view(-1) for s in seeds_d]]*B,[seg_d.view(-1)]*B,out_img=out_img,out_seg=out_seg,scale=True)
def step_two(split=4):
    ids=step_ids(cnt[0],B,0,1); cnt[0]+=1
    pl=plans_for(ids)
    with torch.cuda.stream(sA):
        engA.run_base(pl[:split],[[s.view(-1) for s in seeds_d]]*split,[seg_d.view(-1)]*split,out_img=out_img[:split],out_seg=out_seg[:split],scale=True)
    with torch.cuda.stream(sB):
        engB.run_base(pl[split:],[[s.view(-1) for s in seeds_d]]*(B-split),[seg_d.view(-1)]*(B-split),out_img=out_img[split:],out_seg=out_seg[split:],scale=True)
def timeit(fn,n=30):
    for _ in range(4): fn()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
print("one stream  ms/step", timeit(step_one))
print("two streams ms/step", timeit(step_two))
print("one stream  ms/step", timeit(step_one))
print("two streams ms/step", timeit(step_two))
