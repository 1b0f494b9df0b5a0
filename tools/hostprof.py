"""Host-side profile of the production step (cProfile over the Python that issues it): `python tools/hostprof.py`."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from fetalsyngen_b200.data.packed import PackedSeeds  # noqa: E402
from fetalsyngen_b200.sharding import step_ids  # noqa: E402


def main():
    shape, dev, B = (256, 256, 256), "cuda:0", 8
    gen = bench.build_generator(shape, dev)
    subj = [(torch.from_numpy(seg).to(dev), PackedSeeds(words, counts, device=dev)) for _, seg, words, counts in bench.load_subjects(shape)]
    out_img = torch.empty((B, *shape), dtype=torch.float32, device=dev)
    out_seg = torch.empty((B, *shape), dtype=torch.uint8, device=dev)

    def step(k):
        ids = step_ids(k, B, 0, 1)
        gen.sample_batch([subj[i % 3][0] for i in ids], [subj[i % 3][1] for i in ids], scale=True, out_img=out_img, out_seg=out_seg, sample_ids=ids, base_seed=1234)

    for k in range(5):
        step(k)
    torch.cuda.synchronize()
    host = 0.0
    for r in range(12):  # rounds of 4 steps from an idle device: the parameter ring lets the host run 8 steps ahead at most
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(4):
            step(4 * r + k)
        host += time.perf_counter() - t0
    print(f"host issue time: {host / 48 * 1000:.3f} ms per step")
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for k in range(50):
        step(k)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
    pstats.Stats(pr).sort_stats("tottime").print_stats(25)


if __name__ == "__main__":
    main()
