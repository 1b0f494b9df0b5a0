"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For each case it (1) down-samples the bundled sub-sta30 segmentation + seeds by nearest index
scaling, (2) writes them as temporary NIfTI files, (3) calls the reference's own
``FetalSynthGen.sample`` on CPU with numpy/torch seeded, while recording every random tensor
(``torch.rand/randn``), every sampled parameter and every stage output, and (4) stores the
result as ``base_<name>.npz``.  Nothing here is imported by the product.
"""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

import ref_import  # noqa: E402
from fetalsyngen_b200.utils.nifti import write_nifti  # noqa: E402

OUT = Path(__file__).resolve().parent


def shrink(vol: np.ndarray, shape) -> np.ndarray:
    idx = [np.floor((np.arange(s) + 0.5) * vol.shape[a] / s).astype(int) for a, s in enumerate(shape)]
    return np.ascontiguousarray(vol[np.ix_(*idx)])


BIG = 1 << 16  # draws with at least this many elements are "volume noise"


def seeded_normal(seed: int, shape) -> np.ndarray:
    """Volume noise of the full-size cases: numpy's legacy MT19937 normals are identical on every host and
    version, so only the seed is committed and the tests regenerate the tensor (a 256^3 draw is 64 MiB)."""
    return np.random.RandomState(int(seed)).standard_normal(int(np.prod(shape))).astype(np.float32).reshape(shape)


class Recorder:
    """Monkeypatches torch.rand/randn so every drawn tensor is logged in call order.  With
    ``np_seed_base`` the volume-sized normal draws are *replaced* by ``seeded_normal(np_seed_base + k)``
    (k = index among the big draws) and only ("randn_np", seed, shape) is logged."""

    def __init__(self, np_seed_base=None):
        self.log = []
        self.np_seed_base = np_seed_base
        self._big = 0

    def __enter__(self):
        self._rand, self._randn = torch.rand, torch.randn

        def rand(*a, **k):
            t = self._rand(*a, **k)
            self.log.append(("rand", t.detach().cpu().numpy().copy()))
            return t

        def randn(*a, **k):
            t = self._randn(*a, **k)
            if self.np_seed_base is not None and t.numel() >= BIG:
                seed = self.np_seed_base + self._big
                self._big += 1
                t = torch.from_numpy(seeded_normal(seed, tuple(t.shape))).to(t.device)
                self.log.append(("randn_np", (seed, tuple(t.shape))))
                return t
            self.log.append(("randn", t.detach().cpu().numpy().copy()))
            return t

        torch.rand, torch.randn = rand, randn
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randn = self._rand, self._randn


def run_case(name, shape, seed, overrides=None, force=None, keep_stages=True, subject="sub-sta30", compact=False):
    """compact (full-size cases): volume noise is seeded (see ``seeded_normal``), the inputs are the committed
    subject fixture (tests/golden/subjects), and only the warped segmentation, a strided sample of the image
    and per-plane sums are stored."""
    ref_import.load_reference()
    import fetalsyngen.generator.augmentation.synthseg as ss
    import fetalsyngen.generator.deformation.affine_nonrigid as an

    cfg = ref_import.reference_generator_config(device="cpu", shape=shape)
    for k, v in (overrides or {}).items():
        node = cfg
        ks = k.split(".")
        for kk in ks[:-1]:
            node = node[kk]
        node[ks[-1]] = v
    gen = ref_import.instantiate(cfg)

    seg_full = ref_import.read_nifti(ref_import.seg_path(subject))[0]
    seg = shrink(seg_full, shape).astype(np.float32)
    tmp = Path(tempfile.mkdtemp(prefix="fsg_golden_"))
    seeds, seed_vols = {}, {}
    for n, d in ref_import.seed_paths(subject).items():
        seeds[n] = {}
        for m, p in d.items():
            if compact and tuple(shape) == tuple(seg_full.shape):
                seeds[n][m] = p  # full size: the reference reads its own bundled files
                continue
            v = shrink(ref_import.read_nifti(p)[0], shape)
            fp = tmp / f"s{n}_m{m}.nii.gz"
            write_nifti(fp, v)
            seeds[n][m] = fp
            seed_vols[(n, m)] = v

    if force:
        for path, val in force.items():
            obj = gen
            ks = path.split(".")
            for kk in ks[:-1]:
                obj = getattr(obj, kk)
            setattr(obj, ks[-1], val)

    cap = {}
    # --- stage hooks (outputs only; arithmetic untouched)
    orig_si = gen.intensity_generator.sample_intensities
    orig_ls = gen.intensity_generator.load_seeds

    def ls(*a, **k):
        out = orig_ls(*a, **k)
        cap["labels"] = out[0].numpy().astype(np.uint8)
        cap["mlabel2subclusters"] = np.array([out[1]["mlabel2subclusters"][m] for m in range(1, 5)])
        return out

    def si(*a, **k):
        out = orig_si(*a, **k)
        cap["intensity"] = out[0].numpy().copy()
        cap["mus"], cap["sigmas"] = out[1]["mus"].numpy().copy(), out[1]["sigmas"].numpy().copy()
        return out

    gen.intensity_generator.load_seeds, gen.intensity_generator.sample_intensities = ls, si

    sd = gen.spatial_deform
    orig_di, orig_ad = sd.deform_image, sd.apply_deformation_and_flip

    def di(shp, A, c2, F):
        cap["A"], cap["c2"] = A.numpy().copy(), c2.numpy().copy()
        if F is not None:
            cap["F"] = F.numpy().copy()
        return orig_di(shp, A, c2, F)

    def ad(image, segmentation, output, xx2, yy2, zz2, flip):
        cap["flip"] = bool(flip)
        if xx2 is not None:
            cap["coords"] = np.stack([xx2.numpy(), yy2.numpy(), zz2.numpy()])
        return orig_ad(image, segmentation, output, xx2, yy2, zz2, flip)

    sd.deform_image, sd.apply_deformation_and_flip = di, ad

    orig_blur = ss.gaussian_blur_3d

    def blur(inp, stds, device):
        cap["stds"] = np.array(stds, dtype=np.float64)
        out = orig_blur(inp, stds, device)
        cap["blurred"] = out.numpy().copy()
        return out

    ss.gaussian_blur_3d = blur

    def wrap(obj, attr, key):
        orig = getattr(obj, attr)

        def w(*a, **k):
            out = orig(*a, **k)
            t = out[0] if isinstance(out, tuple) else out
            cap[key] = t.numpy().copy()
            return out

        setattr(obj, attr, w)

    for attr, key in (("gamma", "gamma_out"), ("biasfield", "bias_out"), ("resampled", "lowres"), ("noise", "noisy")):
        inner = getattr(gen, attr)

        def make(inner=inner, key=key):
            def w(*a, **k):
                out = inner(*a, **k)
                cap[key] = out[0].numpy().copy()
                if key == "lowres":
                    cap["factors"] = None if out[1] is None else np.array(out[1], dtype=np.float64)
                return out

            w.resize_back = getattr(inner, "resize_back", None)
            return w

        setattr(gen, attr, make())

    np.random.seed(seed)
    torch.manual_seed(seed)
    try:
        with Recorder(np_seed_base=100000 * seed if compact else None) as rec:
            out, seg_out, _, params = gen.sample(image=None, segmentation=torch.from_numpy(seg), seeds=seeds)
    finally:
        ss.gaussian_blur_3d = orig_blur
    final = out.numpy().copy()

    # ---- identify the random tensors by call order (Appendix A.3 / A.5 of SURVEY.md)
    log = rec.log
    if compact:
        return _save_compact(name, shape, cfg, subject, log, cap, params, seg_out, final)
    d = {
        "shape": np.array(shape),
        "resolution": np.array(cfg["resolution"], dtype=np.float64),
        "seg_in": seg.astype(np.uint8),
        "labels": cap["labels"],
        "mlabel2subclusters": cap["mlabel2subclusters"],
        "seed_labels": np.array(cfg["intensity_generator"]["seed_labels"]),
        "generation_classes": np.array(cfg["intensity_generator"]["generation_classes"]),
        "mus_u": log[0][1],
        "sigmas_u": log[1][1],
        "mus_perturb": log[2][1],
        "gmm_noise": log[3][1],
        "mus": cap["mus"],
        "sigmas": cap["sigmas"],
        "seg_out": seg_out.numpy().astype(np.uint8),
        "final": final,
        "scaled": ref_import.sys.modules["monai.transforms"].ScaleIntensity(0, 1)(out).numpy().copy(),
    }
    i = 4
    dp = params["deform_params"]
    d["flip"] = np.array(bool(dp["flip"]))
    if dp["affine"] is not None:
        d["rotations"], d["shears"], d["scalings"] = (np.asarray(dp["affine"][k], dtype=np.float64) for k in ("rotations", "shears", "scalings"))
        d["A"], d["c2"] = cap["A"], cap["c2"]
        d["c2_u"] = log[i][1]
        i += 1
        if dp["non_rigid"]:
            d["nonlin_std"] = np.float64(dp["non_rigid"]["nonlin_std"])
            d["size_F_small"] = np.array(dp["non_rigid"]["size_F_small"])
            d["Fsmall_n"] = log[i][1]  # standard normal draw; Fsmall = nonlin_std * this
            i += 1
            if keep_stages:
                d["F"] = cap["F"]
        if keep_stages:
            d["coords"] = cap["coords"]
    if params["gamma_params"]["gamma"] is not None:
        d["gamma"] = np.float64(params["gamma_params"]["gamma"])
    if params["bf_params"]["bf_std"] is not None:
        d["bf_std"] = np.asarray(params["bf_params"]["bf_std"], dtype=np.float64)
        d["bf_size"] = np.array(params["bf_params"]["bf_size"])
        d["bf_n"] = log[i][1]
        i += 1
    if params["resample_params"]["spacing"] is not None:
        d["spacing"] = np.array(params["resample_params"]["spacing"], dtype=np.float64)
        d["stds"] = cap["stds"]
        d["factors"] = cap["factors"]
        if keep_stages:
            d["blurred"] = cap["blurred"]
    if params["noise_params"]["noise_std"] is not None:
        d["noise_std"] = np.float64(params["noise_params"]["noise_std"])
        d["noise"] = log[i][1]
        i += 1
    assert i == len(log), (i, len(log), [(k, v.shape) for k, v in log])
    if keep_stages:
        for k in ("intensity", "gamma_out", "bias_out", "lowres", "noisy"):
            d[k] = cap[k]
    # the four seed volumes that were summed (for the seed-cache / in-kernel sum path)
    for m in range(1, 5):
        d[f"seed_m{m}"] = seed_vols[(int(cap["mlabel2subclusters"][m - 1]), m)].astype(np.int8)
    np.savez_compressed(OUT / f"base_{name}.npz", **d)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items() if k in ("final", "lowres", "Fsmall_n", "bf_n", "spacing", "flip", "gamma")})
    return d


SUB = (slice(1, None, 4), slice(2, None, 4), slice(3, None, 4))  # strided image sample of the compact cases


def _save_compact(name, shape, cfg, subject, log, cap, params, seg_out, final):
    assert log[3][0] == "randn_np", log[3][0]
    d = {
        "shape": np.array(shape), "resolution": np.array(cfg["resolution"], dtype=np.float64), "subject": np.array(subject),
        "mlabel2subclusters": cap["mlabel2subclusters"], "mus": cap["mus"], "sigmas": cap["sigmas"],
        "gmm_noise_seed": np.int64(log[3][1][0]),
        "seg_out": seg_out.numpy().astype(np.uint8),
        "final_sub": np.ascontiguousarray(final[SUB]), "final_min": np.float32(final.min()), "final_max": np.float32(final.max()),
        "final_plane_sums": final.astype(np.float64).sum(axis=(1, 2)), "intensity_plane_sums": cap["intensity"].astype(np.float64).sum(axis=(1, 2)),
    }
    i = 4
    dp = params["deform_params"]
    d["flip"] = np.array(bool(dp["flip"]))
    if dp["affine"] is not None:
        d["A"], d["c2"] = cap["A"], cap["c2"]
        i += 1
        if dp["non_rigid"]:
            d["nonlin_std"] = np.float64(dp["non_rigid"]["nonlin_std"])
            d["Fsmall_n"] = log[i][1]
            i += 1
    if params["gamma_params"]["gamma"] is not None:
        d["gamma"] = np.float64(params["gamma_params"]["gamma"])
    if params["bf_params"]["bf_std"] is not None:
        d["bf_std"] = np.asarray(params["bf_params"]["bf_std"], dtype=np.float64)
        d["bf_n"] = log[i][1]
        i += 1
    if params["resample_params"]["spacing"] is not None:
        d["spacing"] = np.array(params["resample_params"]["spacing"], dtype=np.float64)
        d["stds"] = cap["stds"]
        d["factors"] = cap["factors"]
        d["lowres_shape"] = np.array(cap["lowres"].shape)
    if params["noise_params"]["noise_std"] is not None:
        d["noise_std"] = np.float64(params["noise_params"]["noise_std"])
        kind, val = log[i]
        assert kind == "randn_np", kind
        d["noise_seed"], d["noise_shape"] = np.int64(val[0]), np.array(val[1])
        i += 1
    assert i == len(log), (i, len(log), [(k, getattr(v, "shape", v)) for k, v in log])
    np.savez_compressed(OUT / f"full_{name}.npz", **d)
    print(name, {k: (v.shape if getattr(v, "shape", ()) else v) for k, v in d.items() if k not in ("seg_out", "final_sub", "Fsmall_n")})
    return d


SMALL = {
    "spatial_deform.nonlin_scale_min": 0.10,
    "spatial_deform.nonlin_scale_max": 0.20,
    "bias_field.scale_min": 0.05,
    "bias_field.scale_max": 0.15,
}
ALL_ON = {"spatial_deform.prob": 1.0, "resampled.prob": 1.0, "biasfield.prob": 1.0, "gamma.prob": 1.0, "noise.prob": 1.0}

def full_cases():
    """BASELINE.json configs[0]: the bundled 256^3 subjects at the default configuration (stage gates forced)."""
    run_case("sta30_all_flip", (256, 256, 256), 1234, None, {**ALL_ON, "spatial_deform.flip_prb": 1.0}, subject="sub-sta30", compact=True)
    run_case("sta38_noresample", (256, 256, 256), 77, None, {**ALL_ON, "resampled.prob": 0.0, "spatial_deform.flip_prb": 0.0}, subject="sub-sta38", compact=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "full":
        full_cases()
        raise SystemExit(0)
    run_case("c32_all", (32, 32, 32), 1234, SMALL, ALL_ON)
    run_case("c32_flip", (32, 32, 32), 7, SMALL, {**ALL_ON, "spatial_deform.flip_prb": 1.0})
    run_case("c32_noflip", (32, 32, 32), 8, SMALL, {**ALL_ON, "spatial_deform.flip_prb": 0.0})
    run_case("c32_affine_only", (32, 32, 32), 11, {**SMALL, "spatial_deform.nonlinear_transform": False}, ALL_ON)
    run_case("c32_gates_off", (32, 32, 32), 5, SMALL, {"spatial_deform.prob": 0.0, "resampled.prob": 0.0, "biasfield.prob": 0.0, "gamma.prob": 0.0, "noise.prob": 0.0})
    run_case("c32_noresample", (32, 32, 32), 21, SMALL, {**ALL_ON, "resampled.prob": 0.0})
    run_case("c_noncubic", (40, 48, 36), 99, SMALL, ALL_ON)
    run_case("c64_default", (64, 64, 64), 1234, None, ALL_ON, keep_stages=False)
