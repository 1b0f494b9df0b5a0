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


# ------------------------------------------------------------------ full-size (256^3) cases: BASELINE.json configs[0]
SUBJECTS = GOLDEN / "subjects"
FULL_CASES = sorted(p.stem[len("full_"):] for p in GOLDEN.glob("full_*.npz"))
SUB = (slice(1, None, 4), slice(2, None, 4), slice(3, None, 4))  # strided image sample stored by make_golden.py full


def seeded_normal(seed: int, shape) -> np.ndarray:
    """Volume noise of the full-size cases (same function as tests/golden/make_golden.py): numpy's legacy
    MT19937 normals are identical on every host, so the goldens carry only the seed."""
    return np.random.RandomState(int(seed)).standard_normal(int(np.prod(shape))).astype(np.float32).reshape(tuple(int(v) for v in shape))


def load_subject(name):
    """(uint8 segmentation, packed seed words, counts) of a committed subject fixture (the reference's bundled
    sub-sta21/30/38 at 256^3, bit-packed by fetalsyngen_b200/data/packed.py)."""
    with np.load(SUBJECTS / f"{name}.fsgpack.npz") as z:
        return z["seg"], z["words"], z["counts"].tolist()


def load_full_case(name):
    """Golden of the unmodified reference at 256^3 + its inputs: returns (d, labels, seg_in, oracle params)
    with the seeded volume noise regenerated."""
    from fetalsyngen_b200.data.packed import unpack_numpy

    with np.load(GOLDEN / f"full_{name}.npz", allow_pickle=False) as z:
        d = {k: z[k] for k in z.files}
    seg_in, words, counts = load_subject(str(d["subject"]))
    m2s = {m: int(d["mlabel2subclusters"][m - 1]) for m in range(1, 5)}
    labels = unpack_numpy(words, counts, m2s)
    d["shape"] = np.array(seg_in.shape)
    d["gmm_noise"] = seeded_normal(d["gmm_noise_seed"], seg_in.shape)
    if "noise_seed" in d:
        d["noise"] = seeded_normal(d["noise_seed"], d["noise_shape"])
    return d, labels, seg_in, oracle_params(d)
