"""Shared helpers for the -m gpu parity tests (CUDA path vs oracle / golden vectors)."""
import numpy as np
import torch

from fetalsyngen_b200.engine import SamplePlan, engine_for
from golden_util import load_case

DEV = "cuda:0"
TOL = 1e-4  # max-abs error relative to the intensity range (north star)


def rel_err(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    rng = float(b.max() - b.min()) or 1.0
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) / rng


def plan_from_golden(d, dev=DEV):
    shape = tuple(int(v) for v in d["shape"])
    p = SamplePlan(mus=d["mus"], sigmas=d["sigmas"])
    p.gmm_noise = torch.from_numpy(d["gmm_noise"]).to(dev).contiguous().view(-1)
    p.flip = bool(d["flip"])
    if "A" in d:
        p.deform, p.A, p.c2 = True, d["A"], d["c2"]
        p.center = ((np.array(shape) - 1) / 2).astype(np.float32)
        if "Fsmall_n" in d:
            p.fsmall = (np.float32(d["nonlin_std"]) * d["Fsmall_n"]).astype(np.float32)
    if "gamma" in d:
        p.gamma = float(d["gamma"])
    if "bf_n" in d:
        p.bf_low = (d["bf_std"].astype(np.float32) * d["bf_n"]).astype(np.float32)
    if "spacing" in d:
        p.spacing, p.stds = d["spacing"], d["stds"]
    if "noise_std" in d:
        p.noise_std = float(d["noise_std"])
        p.noise = torch.from_numpy(d["noise"]).to(dev).contiguous().view(-1)
    return p


def seeds_from_golden(d, dev=DEV):
    return [torch.from_numpy(d[f"seed_m{m}"]).to(dev).contiguous().view(-1) for m in range(1, 5)]


def engine_from_golden(d, dev=DEV):
    return engine_for(dev, tuple(int(v) for v in d["shape"]), tuple(float(r) for r in d["resolution"]))
