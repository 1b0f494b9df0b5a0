"""One-off parity fuzz on the GPU box: random tile-aligned (texture hand-over) and ragged (generic kernels) shapes,
random draws and random gates (deformation on / off, resolution simulation on / off) — the CUDA base path against the
numpy oracle, segmentation bit for bit and image within the tolerance (the body of
tests/test_gpu_base.py::test_base_pipeline_vs_oracle_random over many more cases).  `python tools/fuzz_parity.py [n]`."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "oracle"), str(ROOT / "tests")]
import np_oracle as O  # noqa: E402
import test_gpu_base as T  # noqa: E402
from fetalsyngen_b200.engine import engine_for  # noqa: E402
from gpu_util import DEV, TOL, rel_err  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rs0 = np.random.RandomState(2026)
    worst = 0.0
    for case in range(n):
        if case % 2 == 0:
            shape = (int(8 * rs0.randint(3, 10)), int(4 * rs0.randint(5, 20)), int(4 * rs0.randint(5, 20)))  # fast path, texture hand-over
        else:
            shape = tuple(int(rs0.randint(17, 80)) for _ in range(3))  # mostly ragged: generic kernels, linear hand-over
        deform, resample = bool(rs0.rand() < 0.8), bool(rs0.rand() < 0.75)
        rs = np.random.RandomState(1000 + case)
        seg, seeds = T._phantom(rs, shape)
        p = T._random_plan(rs, shape, DEV, deform=deform, resample=resample)
        eng = engine_for(DEV, shape, (0.5, 0.5, 0.5))
        q = T._plan_to_oracle(p)
        if resample:
            q["noise"] = p.noise.cpu().numpy().reshape(eng.lowres_shape(p.spacing))
        q["gmm_noise"] = q["gmm_noise"].reshape(shape)
        want_img, want_seg, _ = O.generate_base(sum(s.astype(np.int64) for s in seeds), seg, q)
        dseeds = [torch.from_numpy(s).to(DEV).view(-1) for s in seeds]
        img, sg = eng.run_base([p], [dseeds], [torch.from_numpy(seg).to(DEV).view(-1)])
        assert np.array_equal(sg[0].cpu().numpy(), want_seg), (case, shape, deform, resample, "segmentation")
        err = rel_err(img[0], want_img)
        assert err <= TOL, (case, shape, deform, resample, err)
        worst = max(worst, err)
        print(case, shape, "deform" if deform else "-", "resample" if resample else "-", f"{err:.2e}", flush=True)
    print(f"all {n} cases passed; worst image error / range {worst:.2e} (tolerance {TOL:.0e})")


if __name__ == "__main__":
    main()
