import sys, time, cProfile, pstats, io; sys.path.insert(0,'.')
import numpy as np, torch
import bench
from fetalsyngen_b200.engine import SamplePlan
from fetalsyngen_b200.sharding import sample_seed
gen = bench.build_generator((256,256,256), "cuda:0")
shape=(256,256,256)
def one(i):
    sd32 = sample_seed(1234, i); np.random.seed(sd32); torch.default_generator.manual_seed(sd32)
    plan = SamplePlan(rng_seed=1234, sample_id=i)
    pr = gen._draw_generate(plan, None, shape, {}, None, device_grids=True)
    plan.mus, plan.sigmas = gen.intensity_generator.draw_gmm({})
    pr.update(gen._draw_augment(plan, shape, {}, None, device_grids=True))
    return plan, pr
for i in range(50): one(i)
t0=time.perf_counter()
for i in range(2000): one(i)
print("per sample us", (time.perf_counter()-t0)/2000*1e6)
pr=cProfile.Profile(); pr.enable()
for i in range(2000): one(i)
pr.disable(); s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats('tottime').print_stats(22); print(s.getvalue()[:4500])
