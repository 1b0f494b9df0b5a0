"""Golden vectors for the SimulateMotion artifact, produced by running the UNMODIFIED reference
``Scanner.scan`` + ``PSFReconstructor.recon_psf`` (``fetalsyngen/generator/artifacts/simulate_reco.py``)
on CPU in the build container.

    python tests/golden/make_golden_motion.py

The reference's native extension only runs on a GPU and its CPU fallback (sparse matrices,
slice_acq.py:266-546) is *not* numerically equivalent to it, so the two
dispatchers ``slice_acquisition`` / ``slice_acquisition_adjoint`` are pointed at the numpy
restatement of the CUDA kernels (``oracle/np_motion.py``), which ``tests/test_gpu_motion.py`` pins
against the reference's own extension (``oracle/_ref``) on the GPU box.  Everything else —
draw order, transforms, PSFs, stack selection, slice artifacts, mis-registration, slice removal,
smoothing, Perlin merge — is the reference's own code.  Every random tensor is recorded.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

import np_motion as M  # noqa: E402
import ref_import  # noqa: E402

OUT = Path(__file__).resolve().parent


def inputs(n=48):
    d = np.load(OUT / "base_c64_default.npz")
    o = (64 - n) // 2
    img = d["final"].astype(np.float32)[o : o + n, o : o + n, o : o + n]
    seg = d["seg_out"].astype(np.float32)[o : o + n, o : o + n, o : o + n]
    return np.ascontiguousarray(img), np.ascontiguousarray(seg)


def run(seed, name, scanner_over=None, recon_over=None, n=48):
    ref_import.load_reference()
    import fetalsyngen.generator.artifacts.simulate_reco as SR
    import fetalsyngen.generator.artifacts.utils as U
    from dataclasses import asdict, fields

    img, seg = inputs(n)
    cap = {"fwd": [], "adj": []}

    def fwd(transforms, vol, vol_mask, slices_mask, psf, slice_shape, res_slice, need_weight, interp_psf):
        assert vol_mask is None and slices_mask is None and not need_weight and not interp_psf
        out = M.slice_acq_forward(transforms.numpy(), vol.numpy()[0, 0], psf.numpy(), slice_shape, res_slice)
        cap["fwd"].append({"mat": transforms.numpy().copy(), "psf": psf.numpy().copy(), "res_slice": float(res_slice), "out": out})
        return torch.from_numpy(out)[:, None]

    def adj(transforms, psf, slices, slices_mask, vol_mask, vol_shape, res_slice, interp_psf, equalize):
        assert vol_mask is None and slices_mask is None and interp_psf and equalize
        vol, _ = M.slice_acq_adjoint(transforms.numpy(), psf.numpy(), slices.numpy()[:, 0], tuple(vol_shape), res_slice, True)
        cap["adj"].append({"mat": transforms.numpy().copy(), "psf": psf.numpy().copy(), "slices": slices.numpy()[:, 0].copy(), "res_slice": float(res_slice), "out": vol})
        return torch.from_numpy(vol)[None, None]

    orig = (SR.slice_acquisition, SR.slice_acquisition_adjoint)
    SR.slice_acquisition, SR.slice_acquisition_adjoint = fwd, adj

    sp = dict(resolution_slice_fac_min=0.5, resolution_slice_fac_max=2, resolution_slice_max=1.5, slice_thickness_min=1.5, slice_thickness_max=3.5, gap_min=1.5, gap_max=5.5,
              min_num_stack=2, max_num_stack=6, max_num_slices=250, noise_sigma_min=0, noise_sigma_max=0.1, TR_min=1, TR_max=2, prob_void=0.2, prob_gamma=0.1, gamma_std=0.05,
              slice_size=None, restrict_transform=False, txy=3.0)
    sp.update(scanner_over or {})
    mp = U.ReconMergeParams(merge_type="perlin", perlin_res_list=[1, 2], perlin_octaves_list=[1, 2, 4], perlin_persistence=0.5, perlin_lacunarity=2,
                            gauss_ngaussians_min=2, gauss_ngaussians_max=4, perlin_increase_size=0.25)
    rp = dict(prob_misreg_slice=0.1, slices_misreg_ratio=0.1, prob_misreg_stack=0.1, txy=3.0, prob_merge=1.0, merge_params=mp, prob_smooth=0.2, prob_rm_slices=0.3,
              rm_slices_min=0.1, rm_slices_max=0.4)
    rp.update(recon_over or {})
    scanner_args, recon_args = U.ScannerParams(**sp), U.ReconParams(**rp)
    res = [0.5, 0.5, 0.5]
    scanner_args.resolution_recon = np.float64(res[0])
    scanner = SR.Scanner(**asdict(scanner_args))
    recon = SR.PSFReconstructor(**{f.name: getattr(recon_args, f.name) for f in fields(recon_args)})

    # --- record the per-stack torch draws of add_noise / signal_void
    stacks_log = []
    o_noise, o_void, o_gamma = scanner.add_noise, scanner.signal_void, scanner.random_gamma
    tlog = {"randn_like": [], "rand": []}
    o_randn_like, o_rand, o_rand_like, o_randperm = torch.randn_like, torch.rand, torch.rand_like, torch.randperm
    perms = []

    def randn_like(*a, **k):
        t = o_randn_like(*a, **k)
        tlog["randn_like"].append(t.numpy().copy())
        return t

    def rand(*a, **k):
        t = o_rand(*a, **k)
        tlog["rand"].append(t.numpy().copy())
        return t

    def rand_like(*a, **k):
        t = o_rand_like(*a, **k)
        tlog["rand"].append(t.numpy().copy())
        return t

    def randperm(*a, **k):
        t = o_randperm(*a, **k)
        perms.append(t.numpy().copy())
        return t

    def gamma_w(slices, genparams={}):
        before = slices.numpy().copy()
        out = o_gamma(slices, genparams)
        stacks_log.append({"gamma_applied": np.bool_(not np.array_equal(before, out.numpy()))})
        return out

    def noise_w(slices, genparams={}):
        mask = (slices > scanner.slice_noise_threshold).numpy()
        tlog["randn_like"].clear()
        out = o_noise(slices, genparams)
        stacks_log[-1].update({"noise_mask": np.packbits(mask[:, 0]), "noise1": tlog["randn_like"][0].copy(), "noise2": tlog["randn_like"][1].copy()})
        return out

    def void_w(slices):
        tlog["rand"].clear()
        h, w = slices.shape[-2:]
        out = o_void(slices)
        r = tlog["rand"]
        idx = np.nonzero(r[0] < scanner.prob_void)[0]
        v = {"idx": idx.astype(np.int32)}
        if len(idx):
            v.update({"yc": (r[1] - 0.5) * (h - 1), "xc": (r[2] - 0.5) * (w - 1), "theta": (2 * np.pi * r[3]).reshape(-1), "a": (30 + r[4] * 90).reshape(-1),
                      "A": (r[5] * 0.5 + 0.5).reshape(-1), "sx": (r[6] * 30 + 39).reshape(-1)})
        stacks_log[-1].update({"void": v})
        return out

    scanner.random_gamma, scanner.add_noise, scanner.signal_void = gamma_w, noise_w, void_w
    torch.randn_like, torch.rand, torch.rand_like, torch.randperm = randn_like, rand, rand_like, randperm
    shape = img.shape
    dshape = (1, 1, *shape)
    timg, tseg = torch.from_numpy(img), torch.from_numpy(seg)
    d = {"resolution": np.float64(res[0]), "volume": timg.view(dshape).float(), "mask": (tseg > 0).view(dshape).float(), "seg": tseg.view(dshape).float(),
         "affine": torch.diag(torch.tensor(res + [1])), "threshold": 0.1}
    np.random.seed(seed)
    torch.manual_seed(seed)
    try:
        d_scan = scanner.scan(d)
        n_scan_perms = len(perms)
        tlog["rand"].clear()
        # generate_fractal_noise_3d reseeds numpy from the wall clock (artifacts/utils.py:365-367): harmless here,
        # the Perlin gradients come from torch.rand and are recorded.
        out, weight = recon.recon_psf(d_scan)
    finally:
        SR.slice_acquisition, SR.slice_acquisition_adjoint = orig
        torch.randn_like, torch.rand, torch.rand_like, torch.randperm = o_randn_like, o_rand, o_rand_like, o_randperm
    assert n_scan_perms == 0
    seeds = recon.get_seeds()
    g = {"seed": np.int64(seed), "image": img, "seg": seg.astype(np.uint8), "output": out.numpy()[0, 0].copy(), "weight": weight.numpy().reshape(shape).copy(),
         "recon": cap["adj"][0]["out"], "adj_mat": cap["adj"][0]["mat"], "adj_nslices": np.int64(cap["adj"][0]["slices"].shape[0]), "psf_rec": cap["adj"][0]["psf"],
         "res_slice": np.float64(cap["adj"][0]["res_slice"]), "psf_acq": cap["fwd"][0]["psf"],
         "resolution_slice": np.float64(d_scan["resolution_slice"]), "slice_thickness": np.float64(d_scan["slice_thickness"]), "gap": np.float64(d_scan["gap"]),
         "positions": d_scan["positions"].numpy(), "stacks": d_scan["stacks"].numpy()[:, 0], "transforms": d_scan["transforms"].numpy(), "transforms_gt": d_scan["transforms_gt"].numpy(),
         "n_attempts": np.int64(len(cap["fwd"]) // 2), "n_stacks_logged": np.int64(len(stacks_log)),
         "smooth_volume_on": np.bool_(seeds["smooth_volume_on"]), "rm_slices_on": np.bool_(seeds["rm_slices_on"]), "misreg_slice_on": np.bool_(seeds["misreg_slice_on"]),
         "misreg_stack_on": np.array(seeds["misreg_stack_on"], dtype=np.bool_), "rm_slices_ratio": np.float64(np.nan if seeds["rm_slices_ratio"] is None else seeds["rm_slices_ratio"]), "res": np.int64(seeds["res"]), "octave": np.int64(seeds["octave"])}
    # forward calls come in (image, mask) pairs, one pair per attempted stack
    for k in range(len(cap["fwd"]) // 2):
        g[f"fwd_mat_{k}"] = cap["fwd"][2 * k]["mat"]
        g[f"fwd_img_{k}"] = cap["fwd"][2 * k]["out"]
        g[f"fwd_mask_sums_{k}"] = cap["fwd"][2 * k + 1]["out"].sum((1, 2))
        if k == 0:
            g["fwd_mask_0"] = cap["fwd"][1]["out"]
    for k, s in enumerate(stacks_log):
        for key in ("gamma_applied", "noise_mask", "noise1", "noise2"):
            g[f"{key}_{k}"] = s[key]
        for key, v in s["void"].items():
            g[f"void_{key}_{k}"] = np.asarray(v)
    pi = 0
    if seeds["misreg_slice_on"]:
        g["perm_misreg"] = perms[pi]
        pi += 1
    if seeds["rm_slices_on"]:
        g["perm_kept"] = perms[pi]
        pi += 1
    assert pi == len(perms), (pi, len(perms))
    r = tlog["rand"]
    assert len(r) == 2 * int(seeds["octave"]), (len(r), seeds)
    for o in range(int(seeds["octave"])):
        g[f"theta_{o}"], g[f"phi_{o}"] = r[2 * o], r[2 * o + 1]
    np.savez_compressed(OUT / f"{name}.npz", **g)
    print(name, "stacks", len(stacks_log), "slices", g["stacks"].shape, "attempts", int(g["n_attempts"]), {k: seeds[k] for k in seeds},
          "voids", [int(len(s["void"]["idx"])) for s in stacks_log], "size", (OUT / f"{name}.npz").stat().st_size)


if __name__ == "__main__":
    run(11, "motion_default", scanner_over={"max_num_stack": 3})
    run(7, "motion_all_on", scanner_over={"prob_gamma": 1.0, "prob_void": 0.5, "min_num_stack": 3, "max_num_stack": 3},
        recon_over={"prob_misreg_slice": 1.0, "prob_misreg_stack": 1.0, "prob_smooth": 1.0, "prob_rm_slices": 1.0})
