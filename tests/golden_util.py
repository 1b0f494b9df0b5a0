"""Helpers shared by the parity tests: load a golden case and turn it into oracle parameters."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
BASE_CASES = sorted(p.stem[len("base_"):] for p in GOLDEN.glob("base_*.npz"))


def load_case(name):
    with np.load(GOLDEN / f"base_{name}.npz", allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def oracle_params(d):
    """The drawn parameters + noise tensors of a golden case, in np_oracle.generate_base form."""
    p = {
        "mus": d["mus"],
        "sigmas": d["sigmas"],
        "gmm_noise": d["gmm_noise"],
        "flip": bool(d["flip"]),
        "resolution": d["resolution"],
        "size": tuple(int(v) for v in d["shape"]),
    }
    if "A" in d:
        p["A"], p["c2"] = d["A"], d["c2"]
        if "Fsmall_n" in d:
            p["Fsmall"] = (np.float32(d["nonlin_std"]) * d["Fsmall_n"]).astype(np.float32)
    if "gamma" in d:
        p["gamma"] = float(d["gamma"])
    if "bf_n" in d:
        p["bf_low"] = (d["bf_std"].astype(np.float32) * d["bf_n"]).astype(np.float32)
    if "spacing" in d:
        p["spacing"], p["stds"] = d["spacing"], d["stds"]
    if "noise_std" in d:
        p["noise_std"], p["noise"] = float(d["noise_std"]), d["noise"]
    return p
