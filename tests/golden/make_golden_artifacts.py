"""Golden vectors for the SR-artifact stages, produced by running the UNMODIFIED reference
classes (``fetalsyngen/generator/augmentation/artifacts.py``) on CPU in the build container.

    python tests/golden/make_golden_artifacts.py

Input image / segmentation are the outputs of the base golden case ``base_c64_default``.
Every random tensor (``torch.rand/randn/randperm/multinomial``), every drawn parameter and the
interesting intermediates (MoG weights, Perlin weight, masks) are recorded so that the oracle
and the CUDA kernels can be driven with exactly the reference's draws.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

import ref_import  # noqa: E402

OUT = Path(__file__).resolve().parent


class TorchLog:
    NAMES = ("rand", "randn", "randperm", "multinomial")

    def __init__(self):
        self.log = []

    def __enter__(self):
        self._orig = {n: getattr(torch, n) for n in self.NAMES}
        for n in self.NAMES:
            def make(n=n):
                def f(*a, **k):
                    t = self._orig[n](*a, **k)
                    self.log.append((n, t.detach().cpu().numpy().copy()))
                    return t
                return f
            setattr(torch, n, make())
        return self

    def __exit__(self, *exc):
        for n, f in self._orig.items():
            setattr(torch, n, f)

    def take(self, name):
        for i, (n, v) in enumerate(self.log):
            if n == name:
                return self.log.pop(i)[1]
        raise KeyError(name)


def base_inputs():
    d = np.load(OUT / "base_c64_default.npz")
    return d["final"].astype(np.float32), d["seg_out"].astype(np.float32)


def wrap_module_fn(mod, name, cap, key, grab_args=None):
    orig = getattr(mod, name)

    def w(*a, **k):
        out = orig(*a, **k)
        rec = {"out": out.detach().cpu().numpy().copy() if torch.is_tensor(out) else out}
        if grab_args:
            rec.update(grab_args(*a, **k))
        cap.setdefault(key, []).append(rec)
        return out

    setattr(mod, name, w)
    return orig


def blur_cortex(seed=3):
    ref_import.load_reference()
    import fetalsyngen.generator.augmentation.artifacts as A

    img, seg = base_inputs()
    cap = {}
    o_mog = wrap_module_fn(A, "mog_3d_tensor", cap, "mog", lambda shape, centers, sigmas, device: {
        "centers": np.array([[float(c) for c in ce] for ce in centers], dtype=np.float64), "sigmas": np.array(sigmas, dtype=np.float64)})
    o_blur = wrap_module_fn(A, "gaussian_blur_3d", cap, "blur", lambda inp, stds, device: {"stds": np.array(stds, dtype=np.float64)})
    art = A.BlurCortex(prob=1.0, cortex_label=2, nblur_min=50, nblur_max=200)
    np.random.seed(seed)
    torch.manual_seed(seed)
    try:
        with TorchLog() as tl:
            out, meta = art(torch.from_numpy(img), torch.from_numpy(seg), "cpu", {})
    finally:
        A.mog_3d_tensor, A.gaussian_blur_3d = o_mog, o_blur
    prior, mog = cap["mog"]
    d = {
        "nblur": np.int64(meta["nblur"]),
        "prior": prior["out"], "prior_centers": prior["centers"], "prior_sigmas": prior["sigmas"],
        "multinomial_idx": tl.take("multinomial"),
        "centers": mog["centers"].astype(np.int64), "sigmas": mog["sigmas"], "gaussian": mog["out"],
        "std_blurs": cap["blur"][0]["stds"], "output": out.numpy().copy(),
    }
    np.savez_compressed(OUT / "art_blur_cortex.npz", **d)
    print("blur_cortex", d["nblur"], d["std_blurs"], float(np.abs(d["output"] - img).max()))


def struct_noise(seed=5, merge="perlin", name="art_struct_noise"):
    ref_import.load_reference()
    import fetalsyngen.generator.augmentation.artifacts as A
    import fetalsyngen.generator.artifacts.utils as U

    img, seg = base_inputs()
    cap = {}
    o_frac = wrap_module_fn(A, "generate_fractal_noise_3d", cap, "fractal")
    o_perlin = wrap_module_fn(U, "generate_perlin_noise_3d", cap, "perlin", lambda shape, res, *a, **k: {"res": np.array([int(r) for r in res])})
    o_interp = torch.nn.functional.interpolate
    interp_out = []

    def interp(*a, **k):
        out = o_interp(*a, **k)
        interp_out.append(out.detach().numpy().copy().squeeze())
        return out

    torch.nn.functional.interpolate = interp
    mp = U.StructNoiseMergeParams(merge_type=merge, gauss_nloc_min=5, gauss_nloc_max=15, gauss_sigma_mu=25, gauss_sigma_std=5,
                                  perlin_res_list=[1, 2], perlin_octaves_list=[1, 2, 4], perlin_persistence=0.5, perlin_lacunarity=2, perlin_increase_size=0.1)
    art = A.StructNoise(prob=1.0, wm_label=3, std_min=0.2, std_max=0.4, merge_params=mp, nstages_min=1, nstages_max=5)
    np.random.seed(seed)
    torch.manual_seed(seed)
    state = np.random.get_state()
    try:
        with TorchLog() as tl:
            out, meta = art(torch.from_numpy(img), torch.from_numpy(seg), "cpu", {})
    finally:
        A.generate_fractal_noise_3d, U.generate_perlin_noise_3d = o_frac, o_perlin
        torch.nn.functional.interpolate = o_interp
    nst = int(meta["nstages"])
    d = {"nstages": np.int64(nst), "noise_std": np.float64(meta["noise_std"]),
         "res": np.int64(meta["res"]), "octave": np.int64(meta["octave"]), "weight": cap["fractal"][0]["out"], "output": out.numpy().copy(),
         "lr_noise": interp_out[-1]}
    for k in range(nst):
        d[f"randn_{k}"] = tl.take("randn")
    for o in range(int(meta["octave"])):
        d[f"theta_{o}"] = tl.take("rand")
        d[f"phi_{o}"] = tl.take("rand")
        d[f"perlin_{o}"] = cap["perlin"][o]["out"]
    assert not tl.log, [(n, v.shape) for n, v in tl.log]
    if int(meta["octave"]) > 1:  # keep the fixture small: first octave + the final weight pin the rest
        for k in [k for k in d if (k.startswith("perlin_") and k != "perlin_0") or k == "lr_noise"]:
            del d[k]
    np.savez_compressed(OUT / f"{name}.npz", **d)
    print(name, {k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in d.items() if k not in ("image", "seg")})


def boundaries(seed=2, name="art_boundaries"):
    ref_import.load_reference()
    import fetalsyngen.generator.augmentation.artifacts as A

    img, seg = base_inputs()
    cap = {}
    o_mog = wrap_module_fn(A, "mog_3d_tensor", cap, "mog", lambda shape, centers, sigmas, device: {
        "centers": np.array([[int(c) for c in ce] for ce in centers], dtype=np.int64), "sigmas": np.array(sigmas, dtype=np.float64)})
    art = A.SimulatedBoundaries(prob_no_mask=0.0, prob_if_mask_halo=1.0, prob_if_mask_fuzzy=1.0)
    o_halo, o_fuzzy = art.build_halo, art.generate_fuzzy_boundaries
    halos, fuzz = [], []

    def halo(mask, radius):
        out = o_halo(mask, radius)
        halos.append((int(radius), out.numpy().astype(np.uint8)))
        return out

    def fuzzy(mask, *a, **k):
        out = o_fuzzy(mask, *a, **k)
        fuzz.append(out.numpy().astype(np.uint8))
        return out

    art.build_halo, art.generate_fuzzy_boundaries = halo, fuzzy
    np.random.seed(seed)
    torch.manual_seed(seed)
    try:
        with TorchLog() as tl:
            out, meta = art(torch.from_numpy(img), torch.from_numpy(seg), "cpu", {})
    finally:
        A.mog_3d_tensor = o_mog
    n_fuzzy = int(art.n_generate_fuzzy)
    d = {"halo_radius": np.int64(art.halo_radius), "n_generate_fuzzy": np.int64(n_fuzzy),
         "n_centers": np.int64(art.n_centers), "base_sigma": np.int64(art.base_sigma), "mask_halo": halos[0][1],
         "centers": cap["mog"][0]["centers"], "sigmas": cap["mog"][0]["sigmas"], "mog": cap["mog"][0]["out"], "output": out.numpy().copy()}
    for i in range(n_fuzzy):
        d[f"fuzzy_{i}"] = fuzz[i]
        d[f"perm_{i}"] = tl.take("randperm")
    d["perm_centers"] = tl.take("randperm")
    assert not tl.log, [(n, v.shape) for n, v in tl.log]
    np.savez_compressed(OUT / f"{name}.npz", **d)
    print(name, {k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in d.items() if k not in ("image", "seg")})


if __name__ == "__main__":
    which = sys.argv[1:] or ["blur_cortex", "struct_noise", "boundaries"]
    if "blur_cortex" in which:
        blur_cortex()
    if "struct_noise" in which:
        struct_noise()
        # a second draw with several octaves / res 2 (sample_seeds() cannot be pinned: scan seeds)
        struct_noise(seed=14, name="art_struct_noise_oct")
    if "boundaries" in which:
        boundaries()
