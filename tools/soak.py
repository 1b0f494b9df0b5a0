"""Soak run of the production step: 6000 steps of 8 volumes on one GPU, wall-clock throughput, host RSS and device
memory before / after (r02f on a B200: 4 931 volumes/s wall, no growth).  `python tools/soak.py`."""
import sys, time, resource
sys.path.insert(0, '.')
import torch, bench
from fetalsyngen_b200.data.packed import PackedSeeds
from fetalsyngen_b200.sharding import step_ids
shape, dev, B = (256,)*3, "cuda:0", 8
gen = bench.build_generator(shape, dev)
subj = [(torch.from_numpy(seg).to(dev), PackedSeeds(words, counts, device=dev)) for _, seg, words, counts in bench.load_subjects(shape)]
out_img = torch.empty((B, *shape), dtype=torch.float32, device=dev); out_seg = torch.empty((B, *shape), dtype=torch.uint8, device=dev)
def step(k):
    ids = step_ids(k, B, 0, 1)
    return gen.sample_batch([subj[i % 3][0] for i in ids], [subj[i % 3][1] for i in ids], scale=True, out_img=out_img, out_seg=out_seg, sample_ids=ids, base_seed=1234)
for k in range(50): step(k)
torch.cuda.synchronize()
r0 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss; m0 = torch.cuda.memory_allocated(); f0 = torch.cuda.mem_get_info()[0]
t0 = time.time()
for k in range(50, 6050): step(k)
torch.cuda.synchronize()
dt = time.time() - t0
r1 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss; m1 = torch.cuda.memory_allocated(); f1 = torch.cuda.mem_get_info()[0]
print(f"6000 steps in {dt:.2f} s = {6000*B/dt:.0f} volumes/s wall; host RSS {r0/1024:.0f} -> {r1/1024:.0f} MiB; torch allocated {m0>>20} -> {m1>>20} MiB; device free {f0>>20} -> {f1>>20} MiB")
assert float(out_img.max()) == 1.0
