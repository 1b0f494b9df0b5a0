"""Per-entry-point device time of the production step (B volumes of S^3, drawn like bench.py draws them):
`python tools/kbench.py [--steps 20] [--batch 8] [--shape 256]`; `FSG_LIB=variant.so` times a variant build.
CUDA events bracket every C-ABI call (fetalsyngen_b200._lib.stats), so host gaps are not counted."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from fetalsyngen_b200 import _lib  # noqa: E402
from fetalsyngen_b200.sharding import step_ids  # noqa: E402
from fetalsyngen_b200.data.packed import PackedSeeds  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--shape", type=int, default=256)
    ap.add_argument("--tag", default=os.environ.get("FSG_LIB", "default"))
    a = ap.parse_args()
    shape = (a.shape,) * 3
    dev = "cuda:0"
    gen = bench.build_generator(shape, dev)
    subj = [(torch.from_numpy(seg).to(dev), PackedSeeds(words, counts, device=dev)) for _, seg, words, counts in bench.load_subjects(shape)]
    out_img = torch.empty((a.batch, *shape), dtype=torch.float32, device=dev)
    out_seg = torch.empty((a.batch, *shape), dtype=torch.uint8, device=dev)

    def step(k):
        ids = step_ids(k, a.batch, 0, 1)
        gen.sample_batch([subj[i % 3][0] for i in ids], [subj[i % 3][1] for i in ids], scale=True, out_img=out_img, out_seg=out_seg, sample_ids=ids, base_seed=1234)

    for k in range(3):
        step(k)
    torch.cuda.synchronize()
    _lib.stats.reset()
    _lib.stats.timing = True
    for k in range(a.steps):
        step(k)
    _lib.stats.timing = False
    per = {k: round(v[1] / v[0], 4) for k, v in _lib.stats.elapsed_ms().items()}
    print(json.dumps({"tag": a.tag, "steps": a.steps, "batch": a.batch, "shape": a.shape, "sum_ms": round(sum(per.values()), 4), "per_call_ms": per}))


if __name__ == "__main__":
    main()
