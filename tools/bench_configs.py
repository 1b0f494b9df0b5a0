#!/usr/bin/env python
"""Secondary measurements of BASELINE.json (one JSON line each, CUDA-event timed):

  --config artifacts   configs[2]: full pipeline with the four SR artifacts forced on at 256^3,
                       per-artifact milliseconds through the reference-shaped call
                       (artifact(output, seg, device, genparams, resolution=...)).
  --config sweep       configs[3]: 128^3 / 256^3 / 384^3 with the fixed mid-range parameters of
                       SURVEY.md section 8(d) C4; the warp, the fused blur+down-sample and the
                       up-sample kernels alone, batch 1 and 8, GB/s of algorithmic bytes against the
                       measured HBM peak.

    python tools/bench_configs.py --config sweep > profiles/r01_sweep.jsonl
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from fetalsyngen_b200 import _lib  # noqa: E402
from fetalsyngen_b200.engine import SamplePlan, engine_for  # noqa: E402
from fetalsyngen_b200.tables import make_affine_matrix, resample_size, resample_stds  # noqa: E402
from fetalsyngen_b200.utils.phantom import label_phantom  # noqa: E402

DEV = "cuda:0"


def timed(fn, reps=5, warm=2):
    """Wall-to-wall device time of fn (CUDA events around the whole call sequence, host gaps included)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def timed_kernels(fn, reps=5, warm=2, before=None):
    """Device time of the C-ABI calls fn makes (CUDA events bracket each call: the host-side job
    building between launches is not counted).  Returns (ms per fn, {entry point: ms per fn})."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    _lib.stats.reset()
    for _ in range(reps):
        if before is not None:
            before()
        torch.cuda.synchronize()
        _lib.stats.timing = True
        fn()
        _lib.stats.timing = False
    per = {k: v[1] / reps for k, v in _lib.stats.elapsed_ms().items()}
    _lib.stats.reset()
    return sum(per.values()), per


def flush_l2(buf=[None]):
    if buf[0] is None:
        buf[0] = torch.empty(256 * 2**20, dtype=torch.uint8, device=DEV)
    buf[0].fill_(1)


def run_artifacts(args):
    shape = (args.shape,) * 3
    seg_h, seeds_h = label_phantom(shape)
    arts = bench.default_artifacts(1.0)
    gen = bench.build_generator(shape, DEV)
    seg_d = torch.from_numpy(seg_h).to(DEV)
    seeds_d = [torch.from_numpy(s).to(DEV) for s in seeds_h]
    np.random.seed(0)
    torch.manual_seed(0)
    img, seg, _ = gen.sample_batch([seg_d], [seeds_d], scale=False, sample_ids=[0], base_seed=1)
    out, segf = img[0], seg[0].float()
    res = {}
    for name, art in arts.items():
        times = []
        for rep in range(args.reps + 1):
            np.random.seed(10 + rep)
            torch.manual_seed(10 + rep)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            y, meta = art(out, segf, DEV, {}, resolution=[0.5, 0.5, 0.5])
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
            assert y.shape == out.shape and bool(torch.isfinite(y).all())
        res[name] = {"ms_mean": float(np.mean(times[1:])), "ms_min": float(np.min(times[1:])), "ms_max": float(np.max(times[1:])), "first_call_ms": times[0]}
    base = timed(lambda: gen.sample_batch([seg_d], [seeds_d], scale=False, sample_ids=[0], base_seed=1), reps=5)
    total = base + sum(v["ms_mean"] for v in res.values())
    print(json.dumps({"config": "configs[2]: full pipeline + 4 SR artifacts forced on", "shape": list(shape), "base_pipeline_ms": base, "artifacts_ms": res,
                      "volumes_per_s_single_stream": 1000.0 / total, "reps": args.reps, "note": "device time of the reference-shaped per-sample calls (batch 1); host draws included"}))


def sweep_plan(S, rs):
    """Fixed mid-range parameters of SURVEY.md 8(d) C4."""
    shape = (S, S, S)
    res = 0.5 * 256 / S
    p = SamplePlan(mus=(25 + 200 * rs.rand(50)).astype(np.float32), sigmas=(5 + 20 * rs.rand(50)).astype(np.float32), rng_seed=1, sample_id=int(rs.randint(1 << 30)))
    rot = np.array([10.0, -7.0, 5.0]) / 180 * np.pi
    p.deform, p.flip = True, False
    p.A = make_affine_matrix(rot, np.array([0.01, -0.01, 0.005]), np.array([1.05, 0.95, 1.0])).astype(np.float32)
    p.c2 = (np.array(shape) - 1) / 2
    p.center = ((np.array(shape) - 1) / 2).astype(np.float32)
    s = int(round(0.045 * S))
    p.fsmall = (2.0 * rs.randn(s, s, s, 3)).astype(np.float32)
    p.gamma = 1.05
    p.bf_low = (0.15 * rs.randn(3, 3, 3)).astype(np.float32)
    p.spacing = np.array([2 * res] * 3)
    p.stds = resample_stds(p.spacing, [res] * 3, 0.5)  # (0.85 + 0.3 * 0.5) = 1.0 -> sigma = 2 ln5 / pi voxels
    p.noise_std = 10.0
    return p, res


def run_sweep(args):
    peak, _ = bench.peaks()
    for S in args.sizes:
        shape = (S, S, S)
        N = S**3
        seg_h, seeds_h = label_phantom(shape)
        seg_d = torch.from_numpy(seg_h).to(DEV).view(-1)
        seeds_d = [torch.from_numpy(s).to(DEV).view(-1) for s in seeds_h]
        for B in args.batches:
            rs = np.random.RandomState(S + B)
            plans, res = [], None
            for _ in range(B):
                p, res = sweep_plan(S, rs)
                plans.append(p)
            eng = engine_for(DEV, shape, (res,) * 3)
            buf = [torch.empty((B, N), dtype=torch.float32, device=DEV) for _ in range(4)]
            oseg = torch.empty((B, N), dtype=torch.uint8, device=DEV)
            eng.gmm(plans, [seeds_d] * B, buf[0])
            n = resample_size(S, res, 2 * res)
            f3 = (n / S) ** 3
            cases = {
                "fsg_gmm (seed sum + GMM + Philox)": (lambda: eng.gmm(plans, [seeds_d] * B, buf[0]), 8),
                "fsg_warp (deform + trilinear + nearest + gamma + bias)": (lambda: eng.warp(plans, buf[0], [seg_d] * B, buf[1], oseg), 10),
                "fsg_sepconv (blur o down-sample + noise)": (lambda: eng.sepconv(plans, buf[1], buf[2], buf[2], buf[3]), 4 * (1 + f3)),
            }
            info = eng.sepconv(plans, buf[1], buf[2], buf[2], buf[3])
            cases["fsg_zoom_minmax + fsg_zoom (up-sample, /max, ScaleIntensity)"] = (
                lambda: eng.zoom([buf[2][b] for b in range(B)], [i[0] for i in info], [1 / i[1] for i in info], [buf[3][b] for b in range(B)], post=2), 4 * (1 + f3))
            cases["whole base pipeline (run_base)"] = (lambda: eng.run_base(plans, [seeds_d] * B, [seg_d] * B, scale=True), 8 + 10 + 4 * (1 + f3) * 2)
            for name, (fn, bytes_per_voxel) in cases.items():
                small = B * N * 4 < 200 * 2**20  # working set below L2: flush between iterations
                ms, per = timed_kernels(fn, reps=args.reps, before=flush_l2 if small else None)
                algo = bytes_per_voxel * N * B
                print(json.dumps({"config": "configs[3] sweep", "shape": S, "batch": B, "kernel": name, "ms": round(ms, 4), "algorithmic_bytes": int(algo),
                                  "GBps": round(algo / ms / 1e6, 1), "frac_of_measured_hbm_peak": round(algo / ms / 1e6 / peak, 4), "volumes_per_s": round(B * 1000 / ms, 1),
                                  "coarse_grid": n, "l2": "flushed between iterations" if small else "working set > L2", "per_entry_ms": {k: round(v, 4) for k, v in per.items()}}))
            del buf, oseg
            torch.cuda.empty_cache()


def run_motion(args):
    """Slice-acquisition forward / PSF-reconstruction adjoint at the default config's sizes: libfsg vs
    the reference's own extension (oracle/_ref, built from the reference's sources for sm_100a)."""
    from fetalsyngen_b200.generator.artifacts import simulate_reco as SR
    from fetalsyngen_b200.generator.artifacts import svort

    ext = bench.reference_motion_extension()  # baseline leg: lives in bench.py, the one measurement file that may touch oracle/
    S = args.shape
    seg_h, _ = label_phantom((S, S, S))
    rs = np.random.RandomState(0)
    vol = torch.from_numpy((seg_h > 0).astype(np.float32) * (0.3 + 0.7 * rs.rand(S, S, S).astype(np.float32))).to(DEV)
    for res_s, thick, gap in ((0.6, 2.5, 3.5), (1.0, 3.5, 1.6), (0.3, 1.5, 5.0)):
        np.random.seed(1)
        psf = svort.get_PSF(res_ratio=(res_s / 0.5, res_s / 0.5, thick / 0.5))
        ss = int(np.ceil(int(np.sqrt(3 * S * S / 2.0) * 0.5 / res_s) / 32.0) * 32)
        ns = int(S * 0.5 / gap) + 2
        init = svort.random_init_stack_transforms(ns, gap, False, 3.0)
        motion = svort.sample_motion(np.arange(ns) * 1.5, True)
        mat = svort.mat_update_resolution(motion.compose(init).matrix(), 0.5, 0.5)
        nstack = max(1, min(6, 250 // ns))
        mats = np.concatenate([mat] * nstack)[:250]
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
        empty = torch.empty(0, device=DEV)
        ours_f = timed(lambda: SR.slice_acquisition(mat, vol, psf, (ss, ss), res_s / 0.5), reps=args.reps)
        pairs = SR.volume_xpairs(vol)
        ours_fp = timed(lambda: SR.slice_acquisition(mat, vol, psf, (ss, ss), res_s / 0.5, pairs=pairs), reps=args.reps)
        t_pairs = timed(lambda: SR.volume_xpairs(vol), reps=args.reps)
        quads = SR.volume_xyquads(vol)
        ours_fq = timed(lambda: SR.slice_acquisition(mat, vol, psf, (ss, ss), res_s / 0.5, pairs=quads), reps=args.reps)
        sl = SR.slice_acquisition(mats, vol, psf, (ss, ss), res_s / 0.5)
        ours_a = timed(lambda: SR.slice_acquisition_adjoint(mats, psf, sl, (S, S, S), res_s / 0.5), reps=args.reps)
        line = {"config": "motion kernels", "shape": S, "res_slice": res_s, "slice_thickness": thick, "gap": gap, "slice_size": ss, "slices_per_stack": ns, "psf_shape": list(psf.shape),
                "psf_taps": int((psf != 0).sum()), "forward_ms_ours": ours_f, "forward_ms_ours_xpairs": ours_fp, "xpairs_build_ms": t_pairs, "forward_ms_ours_xyquads": ours_fq, "adjoint_slices": int(mats.shape[0]), "adjoint_ms_ours": ours_a}
        if ext is not None:
            tm, tp, tms = t(mat), t(psf), t(mats)
            ref_f = timed(lambda: ext.forward(tm, vol[None, None], empty, empty, tp, [ss, ss], float(res_s / 0.5), False, False), reps=args.reps)
            ref_a = timed(lambda: ext.adjoint_forward(tms, tp, sl, empty, empty, [S, S, S], float(res_s / 0.5), True, True), reps=args.reps)
            line.update({"forward_ms_reference_ext": ref_f, "adjoint_ms_reference_ext": ref_a, "forward_speedup": ref_f / ours_f, "adjoint_speedup": ref_a / ours_a})
        print(json.dumps(line))


def run_sample_api(args):
    """The reference's own published figure is `generation_time` of one `FetalSynthDataset.sample`
    call (docs/datasets.md:76,131: 0.5616 s / 0.6192 s on an unspecified GPU; wall clock around
    generator.sample + ScaleIntensity + .cpu(), datasets.py:303-320).  Same call here, on a BIDS tree
    written to a temporary directory from the label phantom."""
    import tempfile
    import time

    from fetalsyngen_b200.data.datasets import FetalSynthDataset
    from fetalsyngen_b200.utils import nifti

    S = args.shape
    shape = (S, S, S)
    seg_h, seeds_h = label_phantom(shape)
    aff = np.diag([0.5, 0.5, 0.5, 1.0])
    with tempfile.TemporaryDirectory() as tmp:
        root = Path(tmp)
        d = root / "bids" / "sub-phantom" / "anat"
        d.mkdir(parents=True)
        nifti.write_nifti(d / "sub-phantom_rec-x_T2w_dseg.nii.gz", seg_h.astype(np.float32), aff)
        for n in range(1, 7):
            sd = root / "seeds" / f"subclasses_{n}" / "sub-phantom" / "anat"
            sd.mkdir(parents=True)
            _, seeds_n = label_phantom(shape, n_sub=(n, n, n, n), seed=n)  # labels 10 m .. 10 m + n - 1, as generate_seeds.py writes them
            for m in range(1, 5):
                nifti.write_nifti(sd / f"sub-phantom_rec-x_T2w_dseg_mlabel_{m}.nii.gz", seeds_n[m - 1], aff)
        res = {}
        for name, arts in (("base stages (default probabilities)", None), ("with the four SR artifacts (default probabilities)", bench.default_artifacts(0.4))):
            gen = bench.build_generator(shape, DEV, artifacts=arts)
            for st in (gen.spatial_deform, gen.resampled, gen.biasfield, gen.noise, gen.gamma):
                st.prob = 0.9  # configs/dataset/generator/default.yaml
            gen.spatial_deform.flip_prb = 0.5
            ds = FetalSynthDataset(str(root / "bids"), gen, str(root / "seeds"), None)
            np.random.seed(0)
            torch.manual_seed(0)
            t0 = time.perf_counter()
            ds.sample(0)
            first = time.perf_counter() - t0
            t0 = time.perf_counter()
            nwarm = 0
            while len(gen.intensity_generator._cache) < 24 and nwarm < 200:  # all 24 seed volumes decoded once (the reference re-reads 4 per sample)
                ds.sample(0)
                nwarm += 1
            warm = time.perf_counter() - t0
            times = []
            for _ in range(args.reps * 4):
                t0 = time.perf_counter()
                out, params = ds.sample(0)
                times.append(time.perf_counter() - t0)
            assert out["image"].shape == (1, *shape) and out["label"].dtype == torch.int64 and out["image"].device.type == "cpu"
            res[name] = {"first_call_s": first, "seed_cache_warmup_calls": nwarm, "seed_cache_warmup_s": warm, "mean_s": float(np.mean(times)), "median_s": float(np.median(times)), "min_s": float(np.min(times)), "max_s": float(np.max(times)),
                         "generation_time_last": params["generation_time"]}
        # bit-packed subject cache (data/packed.py): one-time conversion, then the cold start of a new process
        from fetalsyngen_b200.data.packed import PackedSeeds, load_packed

        gen = bench.build_generator(shape, DEV, artifacts=None)
        t0 = time.perf_counter()
        ds = FetalSynthDataset(str(root / "bids"), gen, str(root / "seeds"), None, packed_cache=str(root / "cache"))
        ds.sample(0)
        convert = time.perf_counter() - t0
        cache_file = root / "cache" / "sub-phantom.fsgpack.npz"
        gen = bench.build_generator(shape, DEV, artifacts=None)
        t0 = time.perf_counter()
        ds = FetalSynthDataset(str(root / "bids"), gen, str(root / "seeds"), None, packed_cache=str(root / "cache"))
        ds.sample(0)
        cold = time.perf_counter() - t0
        times = []
        for _ in range(args.reps * 4):
            t0 = time.perf_counter()
            ds.sample(0)
            times.append(time.perf_counter() - t0)
        _, ps, _ = load_packed(cache_file)
        ps.on(DEV)
        unpack_ms = timed(lambda: ps.labels({1: 3, 2: 6, 3: 2, 4: 5}, DEV), reps=20)
        res["bit-packed subject cache"] = {"convert_and_first_call_s": convert, "cold_start_first_call_s": cold, "median_s": float(np.median(times)), "cache_file_MiB": cache_file.stat().st_size / 2**20,
                                           "nifti_gz_files_replaced": 25, "device_bytes_per_subject_MiB": ps._host.nbytes / 2**20, "device_bytes_unpacked_int8_MiB": 24 * ps._host.size / 2**20,
                                           "unpack_kernel_ms": unpack_ms}
    print(json.dumps({"config": "FetalSynthDataset.sample wall clock (reference: generation_time 0.5616 s / 0.6192 s, docs/datasets.md:76,131)", "shape": list(shape), "calls": args.reps * 4,
                      "results": res, "note": "host tensors out: float32 image (1,S,S,S) + int64 label on the CPU, as the reference returns them; seeds / segmentation decoded once and cached on the device"}))


def seed_subject(S, seed=11):
    """Synthetic subject for seed generation: phantom labels + a T2w-like image (tissue means, texture,
    smooth shading, non-brain tissue around the labelled region, zero background)."""
    shape = (S, S, S)
    seg, _ = label_phantom(shape, seed=5)
    rs = np.random.RandomState(seed)
    base = np.array([0, 900, 300, 450, 820, 520, 330, 480], dtype=np.float32)[seg]
    tex = rs.standard_normal(shape).astype(np.float32)
    image = np.maximum(base + 40 * tex + 60 * np.sin(np.arange(S, dtype=np.float32) / (S / 13))[None, None, :], 0).astype(np.float32)
    g = np.meshgrid(*[np.linspace(-1, 1, s, dtype=np.float32) for s in shape], indexing="ij", sparse=True)
    r = np.sqrt(sum(gi**2 for gi in g))
    outer = (seg == 0) & (r > 0.85)
    ring = (seg == 0) & (r <= 0.85)
    image[outer] = 0
    image[ring] = np.maximum(200 + 70 * tex[ring], 1)
    return image, seg


def run_seeds(args):
    """scripts/generate_seeds.py for one subject, sub-class counts 1..6 (what the
    bundled seeds use): 20 fits x 5 initialisations.  GPU: SeedGenerator.split_labels from host arrays
    (H2D, partition, k-means++, EM, predict, label volumes on the device).  CPU: the unmodified
    scikit-learn call of the reference (GaussianMixture(k, n_init=5, init_params="k-means++")
    .fit_predict on a torch tensor, i.e. float64) on a bounded sample of the same fits."""
    import time
    import warnings

    from fetalsyngen_b200.seeds import SeedGenerator

    S = args.shape
    image, seg = seed_subject(S)
    subs = list(range(1, 7))
    gen = SeedGenerator("feta", DEV, seed=1)
    out = gen.split_labels(image, seg, subs)  # warm-up (module load, allocator)
    torch.cuda.synchronize()
    times = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        out = gen.split_labels(image, seg, subs)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    x, index, counts = gen.partition(image, seg)
    fits = gen.last_fit
    iters = {f"{s}x{m + 1}": f["n_iter"] for (s, m), f in fits.items()}
    comp_evals = sum(counts[m] * s * f["n_iter"] for (s, m), f in fits.items())  # best initialisation only (lower bound on the work done)
    # device-resident inputs
    img_d, seg_d = torch.from_numpy(image).to(DEV), torch.from_numpy(seg).to(DEV)
    t_dev = timed(lambda: gen.split_labels(img_d, seg_d.cpu().numpy(), subs), reps=args.reps)
    line = {"config": "seed generation, one subject, sub-class counts 1..6 (scripts/generate_seeds.py)", "shape": [S, S, S], "voxels_per_meta_label": counts,
            "fits": len(fits), "initialisations": len(fits) * gen.n_init, "gpu_s_host_arrays": float(np.median(times)), "gpu_s_min": float(np.min(times)), "gpu_ms_device_image": t_dev,
            "n_iter_of_best_init": iters, "component_evaluations_best_inits": int(comp_evals), "seed_volumes": sum(len(v) for v in out.values())}
    # CPU: the reference's sklearn call on a bounded sample of the fits
    try:
        from sklearn.mixture import GaussianMixture
    except Exception:  # noqa: BLE001
        GaussianMixture = None
    off = np.concatenate([[0], np.cumsum(counts)])
    cpu = {}
    if GaussianMixture is not None:
        xh = x.cpu()
        for (s, m) in ((2, 0), (3, 2), (6, 1)):
            xt = xh[off[m] : off[m + 1]].reshape(-1, 1)
            np.random.seed(0)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                t0 = time.perf_counter()
                GaussianMixture(n_components=s, n_init=5, init_params="k-means++").fit_predict(xt)
                cpu[f"{s}x{m + 1}"] = {"n": int(xt.shape[0]), "sklearn_s": time.perf_counter() - t0}
        # the same three fits alone on the GPU
        specs = [(x[off[m] : off[m + 1]], s) for (s, m) in ((2, 0), (3, 2), (6, 1))]
        ms = timed(lambda: gen.fit_jobs(specs), reps=args.reps)
        line["cpu_sample"] = {"fits": cpu, "sklearn_s_total": sum(v["sklearn_s"] for v in cpu.values()), "gpu_ms_same_fits": ms, "cores": torch.get_num_threads(),
                              "kind": "reference dependency (scikit-learn, unmodified), float64"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", choices=["artifacts", "sweep", "motion", "sample_api", "seeds"], required=True)
    ap.add_argument("--shape", type=int, default=256)
    ap.add_argument("--sizes", type=int, nargs="+", default=[128, 256, 384])
    ap.add_argument("--batches", type=int, nargs="+", default=[1, 8])
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device")
    _lib.load()
    {"artifacts": run_artifacts, "sweep": run_sweep, "motion": run_motion, "sample_api": run_sample_api, "seeds": run_seeds}[args.config](args)


if __name__ == "__main__":
    main()
