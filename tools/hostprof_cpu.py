"""Host-side cost of the production step WITHOUT a GPU: the engine runs on CPU tensors with every C-ABI call
replaced by a no-op, so only the Python that draws the parameters, fills the job structs and queues the calls is
timed (what `bench.py` reports as host_ms_per_step, minus the driver's launch cost).  `python tools/hostprof_cpu.py
[--profile]`.  A development aid: nothing here is on a product path."""
import argparse
import cProfile
import pstats
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
sys.path.insert(0, str(ROOT / "tests"))
from fetalsyngen_b200.sharding import step_ids  # noqa: E402
from host_mock import FakePacked, Recorder, cpu_engine, install  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--shape", type=int, default=64, help="volume edge (the host cost does not depend on it; small keeps memory low)")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--generic", action="store_true", help="force the per-sample job builder (engine.run_base)")
    a = ap.parse_args()
    shape, B = (a.shape,) * 3, a.batch
    if a.generic:
        import fetalsyngen_b200.batch_step as bs

        bs.run_base_batch = lambda *args, **kw: False
    install(Recorder(keep=False))
    gen = bench.build_generator(shape, "cpu")
    eng = cpu_engine(shape, gen.resolution)
    gen.engine = lambda shp: eng
    rs = np.random.RandomState(0)
    nv = int(np.prod(shape))
    segs = [torch.from_numpy(rs.randint(0, 8, nv).astype(np.uint8)) for _ in range(3)]

    packed = [FakePacked(shape) for _ in range(3)]
    out_img = torch.empty((B, *shape), dtype=torch.float32)
    out_seg = torch.empty((B, *shape), dtype=torch.uint8)

    def step(k):
        ids = step_ids(k, B, 0, 1)
        gen.sample_batch([segs[i % 3] for i in ids], [packed[i % 3] for i in ids], scale=True, out_img=out_img, out_seg=out_seg, sample_ids=ids, base_seed=1234)

    for k in range(20):
        step(k)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        for k in range(a.steps):
            step(k)
        best = min(best, (time.perf_counter() - t0) / a.steps * 1000)
    print(f"host time without the driver: {best:.3f} ms per step (best of 5 x {a.steps})")
    if a.profile:
        pr = cProfile.Profile()
        pr.enable()
        for k in range(a.steps):
            step(k)
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(30)


if __name__ == "__main__":
    main()
