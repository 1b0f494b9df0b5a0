"""``FetalSynthGen`` — same constructor and ``sample / generate / augment`` contract as the
reference generator (``fetalsyngen/generator/model.py:27-276``), re-implemented as a short
sequence of fused sm_100a kernels:

  reference stage (model.py)                      here
  ----------------------------------------------  -----------------------------------------
  load_seeds + sample_intensities  (:119-128)     fsg_gmm   (seed sum + lookup + noise)
  spatial_deform.deform            (:147-152)     fsg_warp_shift + fsg_warp
  gamma, biasfield                 (:183-190)     epilogue of fsg_warp
  resampled (blur + down-sample)   (:193-198)     fsg_blur3d + fsg_resample
  noise                            (:201-203)     epilogue of fsg_resample
  resampled.resize_back            (:206)         fsg_zoom_minmax + fsg_zoom (/max)
  SR artifacts                     (:209-220)     generator/augmentation/artifacts.py

All random *parameters* are drawn on the host in the reference's order (so ``np.random.seed``
reproduces the same scalars as the reference); per-voxel noise comes from counter-based Philox
streams keyed by a seed drawn from torch's generator.  ``inject`` (not in the reference API)
lets the parity tests supply the reference's captured random tensors.
"""
from __future__ import annotations

from typing import Iterable

import numpy as np
import torch

from ..engine import SamplePlan, engine_for
from .augmentation.synthseg import RandBiasField, RandGamma, RandNoise, RandResample
from .deformation.affine_nonrigid import SpatialDeformation
from .intensity.rand_gmm import ImageFromSeeds


# FSG_FAST_STEP=0: issue the batched throughput path through the per-sample job builder (engine.run_base) instead of
# the vectorised one (batch_step.run_base_batch); same launches, same results, more host time.
_FAST_STEP = (__import__("os").environ.get("FSG_FAST_STEP", "1") or "1") != "0"
# FSG_NATIVE_STEP=0: build the jobs of a batched step in Python (numpy, batch_step.run_base_batch) instead of in the
# library (fsg_step_run).
_NATIVE_STEP = (__import__("os").environ.get("FSG_NATIVE_STEP", "1") or "1") != "0"


class FetalSynthGen:
    def __init__(
        self,
        shape: Iterable[int],
        resolution: Iterable[float],
        device: str,
        intensity_generator: ImageFromSeeds,
        spatial_deform: SpatialDeformation,
        resampler: RandResample,
        bias_field: RandBiasField,
        noise: RandNoise,
        gamma: RandGamma,
        blur_cortex=None,
        struct_noise=None,
        simulate_motion=None,
        boundaries=None,
    ):
        self.shape = shape
        self.resolution = resolution
        self.intensity_generator = intensity_generator
        self.spatial_deform = spatial_deform
        self.resampled = resampler
        self.biasfield = bias_field
        self.gamma = gamma
        self.noise = noise
        self.artifacts = {
            "blur_cortex": blur_cortex,
            "struct_noise": struct_noise,
            "simulate_motion": simulate_motion,
            "boundaries": boundaries,
        }
        self.device = device
        self._sample_counter = 0

    # ------------------------------------------------------------------ helpers
    def _validated_genparams(self, d: dict) -> dict:
        """Drop None-valued keys recursively: they are not pinned (model.py:85-92)."""
        if not isinstance(d, dict):
            return d
        return {k: self._validated_genparams(v) for k, v in d.items() if v is not None}

    def engine(self, shape=None):
        eng = engine_for(self.device, tuple(self.shape if shape is None else shape), self.resolution)
        key = (float(getattr(self.resampled, "min_resolution", 0) or 0), float(getattr(self.resampled, "max_resolution", 0) or 0))
        if key[1] > 0 and getattr(eng, "_prewarmed", None) != key:
            # per-device workspace built once (the reference allocates nothing up front and rebuilds its
            # sampling grids for every sample): all resolution-simulation tables of this generator
            for a in sorted(set(eng.shape)):
                eng.tables.prewarm_resample(a, float(eng.resolution[eng.shape.index(a)]), key[0], key[1])
                # zoom tables of the control grids (deformation: nonlin_scale * S nodes, bias: bf_scale * S)
                for lo, hi in ((getattr(self.spatial_deform, "nonlin_scale_min", 0), getattr(self.spatial_deform, "nonlin_scale_max", 0)),
                               (getattr(self.biasfield, "scale_min", 0), getattr(self.biasfield, "scale_max", 0))):
                    if hi and hi > 0:
                        for n in range(max(int(np.floor(lo * a)), 1), min(int(np.ceil(hi * a)) + 1, a) + 1):
                            eng.tables.zoom(n, a / n, a)
            eng._prewarmed = key
        return eng

    def _new_plan(self) -> SamplePlan:
        self._sample_counter += 1
        return SamplePlan(rng_seed=int(torch.randint(0, 2**62, (1,)).item()), sample_id=self._sample_counter)

    # ------------------------------------------------------------------ draws
    def _draw_generate(self, plan, seeds, shape, genparams, inject, device_grids: bool = False):
        ig = self.intensity_generator
        params = {}
        if seeds is not None:
            m2s = ig.draw_subclusters(genparams=genparams.get("selected_seeds", {}))
            plan.meta["seed_vols"] = ig.select_seeds(seeds, m2s, self.engine(shape).device)
            plan.mus, plan.sigmas = ig.draw_gmm(genparams.get("seed_intensities", {}), inject)
            params["selected_seeds"] = {"mlabel2subclusters": m2s}
        else:
            params["selected_seeds"] = {}
            params["seed_intensities"] = {}
        fields, deform_params = self.spatial_deform.draw(shape, genparams.get("deform_params", {}), inject, device_grids=device_grids)
        for k, v in fields.items():
            setattr(plan, k, v)
        params["deform_params"] = deform_params
        return params

    def _draw_augment(self, plan, shape, genparams, inject, device_grids: bool = False):
        plan.gamma = self.gamma.draw(genparams.get("gamma_params", {}))
        bf, bf_params = self.biasfield.draw(shape, genparams.get("bf_params", {}), inject, device_grids=device_grids)
        if isinstance(bf, tuple):
            plan.bf_dev = bf
        else:
            plan.bf_low = bf
        plan.spacing, plan.stds = self.resampled.draw(np.array(self.resolution), genparams.get("resample_params", {}), inject)
        plan.noise_std = self.noise.draw(genparams.get("noise_params", {}))
        return {
            "gamma_params": {"gamma": plan.gamma},
            "bf_params": bf_params,
            "resample_params": {"spacing": None if plan.spacing is None else plan.spacing.tolist()},
            "noise_params": {"noise_std": plan.noise_std},
        }

    @staticmethod
    def _inject_tensors(plan, inject, device):
        if not inject:
            return
        if "gmm_noise" in inject:
            plan.gmm_noise = torch.as_tensor(inject["gmm_noise"], dtype=torch.float32).to(device).contiguous().view(-1)
        if "noise" in inject:
            plan.noise = torch.as_tensor(inject["noise"], dtype=torch.float32).to(device).contiguous().view(-1)

    # ------------------------------------------------------------------ device stages
    def _run_generate(self, eng, plan, image, segmentation, fuse_augment: bool):
        """GMM (or image prior) + warp; with fuse_augment the gamma/bias epilogues ride along."""
        seg_u8 = eng.to_u8(segmentation.to(eng.device))
        src = torch.empty((1, eng.nvox), dtype=torch.float32, device=eng.device)
        img_dev = None if image is None else image.to(eng.device, torch.float32).contiguous()
        if "seed_vols" in plan.meta:
            eng.gmm([plan], [[v.view(-1) for v in plan.meta["seed_vols"]]], src)
        else:
            if img_dev is None:
                raise ValueError("If no seeds are passed, an image must be loaded to be used as intensity prior!")
            # (image - min) / (max - min) * 255 (model.py:138): libfsg reduction + scale with the range
            # pre-divided by 255 (two floats cross the bus; this is the rarely used image-prior path)
            from .. import _lib
            from ..engine import _stream

            mm = torch.empty(2, dtype=torch.float32, device=eng.device)
            _lib.call("fsg_minmax", img_dev.data_ptr(), img_dev.numel(), mm.data_ptr(), _stream())
            lo, hi = (float(v) for v in mm.cpu())
            mm2 = torch.tensor([lo, lo + (hi - lo) / 255.0], dtype=torch.float32).to(eng.device)
            _lib.call("fsg_scale_intensity", img_dev.data_ptr(), src.data_ptr(), img_dev.numel(), mm2.data_ptr(), _stream())
        dst = torch.empty((1, eng.nvox), dtype=torch.float32, device=eng.device)
        dseg = torch.empty((1, eng.nvox), dtype=torch.uint8, device=eng.device)
        dst2 = None if img_dev is None else torch.empty((1, eng.nvox), dtype=torch.float32, device=eng.device)
        eng.warp([plan], src, [seg_u8.view(-1)], dst, dseg, None if img_dev is None else [img_dev.view(-1)], dst2, epilogue=fuse_augment)
        return dst, dseg, dst2

    def _run_augment_tail(self, eng, plan, x):
        """blur + down-sample (+noise) + up-sample (/max), or full-resolution noise."""
        if plan.spacing is not None:
            need = eng.sep_capacity(plan)  # exceeds the volume when the spacing is finer than the resolution
            low = eng.scratch("buf1", 1, numel=max(eng.nvox, need[0], need[1]))
            tmp = eng.scratch("buf0", 1, numel=max(eng.nvox, need[2]))
            info = eng.sepconv([plan], x, low, low, tmp)
            out = torch.empty_like(x)
            eng.zoom([low[0]], [info[0][0]], [1 / info[0][1]], out, post=1)
            return out
        if plan.noise_std is not None:
            out = torch.empty_like(x)
            eng.add_noise([plan], x, out)
            return out
        return x

    def _run_artifacts(self, output, segmentation, genparams):
        artifacts = {}
        for name, artifact in self.artifacts.items():
            if artifact is not None:
                output, metadata = artifact(output, segmentation, self.device, genparams.get("artifact_params", {}), resolution=self.resolution)
                artifacts[name] = metadata
        return output, artifacts

    # ------------------------------------------------------------------ reference API
    def generate(self, image, segmentation, seeds, genparams: dict = {}, inject: dict | None = None):
        """Synthetic deformed image from seeds (or the image) + deformed segmentation
        (model.py:94-159).  Returns (output, segmentation, image, synth_params)."""
        shape = tuple(segmentation.shape)
        eng = self.engine(shape)
        plan = self._new_plan()
        params = self._draw_generate(plan, seeds, shape, genparams, inject)
        self._inject_tensors(plan, inject, eng.device)
        dst, dseg, dst2 = self._run_generate(eng, plan, image, segmentation, fuse_augment=False)
        if seeds is not None:
            params["seed_intensities"] = {"mus": torch.from_numpy(plan.mus).to(eng.device), "sigmas": torch.from_numpy(plan.sigmas).to(eng.device)}
        seg_out = eng.from_u8(dseg.view(shape), segmentation.dtype)
        return dst.view(shape), seg_out, None if dst2 is None else dst2.view(shape), params

    def augment(self, image, segmentation, genparams: dict = {}, inject: dict | None = None):
        """Intensity / resolution augmentations + SR artifacts (model.py:161-229)."""
        shape = tuple(image.shape)
        eng = self.engine(shape)
        plan = self._new_plan()
        params = self._draw_augment(plan, shape, genparams, inject)
        self._inject_tensors(plan, inject, eng.device)
        x = image.to(eng.device, torch.float32).contiguous().view(1, -1)
        if plan.gamma is not None or plan.bf_low is not None:
            y = torch.empty_like(x)
            eng.warp([plan], x, None, y, None)
            x = y
        out = self._run_augment_tail(eng, plan, x).view(shape)
        out, artifacts = self._run_artifacts(out, segmentation, genparams)
        params["artifacts"] = artifacts
        return out, params

    def sample(self, image, segmentation, seeds, genparams: dict = {}, inject: dict | None = None):
        """generate + augment with the elementwise stages fused into the warp (model.py:231-276)."""
        if genparams:
            genparams = self._validated_genparams(genparams)
        shape = tuple(segmentation.shape)
        eng = self.engine(shape)
        plan = self._new_plan()
        params = self._draw_generate(plan, seeds, shape, genparams, inject)
        params.update(self._draw_augment(plan, shape, genparams, inject))
        self._inject_tensors(plan, inject, eng.device)
        dst, dseg, dst2 = self._run_generate(eng, plan, image, segmentation, fuse_augment=True)
        if seeds is not None:
            params["seed_intensities"] = {"mus": torch.from_numpy(plan.mus).to(eng.device), "sigmas": torch.from_numpy(plan.sigmas).to(eng.device)}
        out = self._run_augment_tail(eng, plan, dst).view(shape)
        seg_out = eng.from_u8(dseg.view(shape), segmentation.dtype)
        out, artifacts = self._run_artifacts(out, seg_out, genparams)
        params["artifacts"] = artifacts
        return out, seg_out, None if dst2 is None else dst2.view(shape), params

    def _batch_artifacts(self, img, seg, params, genparams, sample_ids, base_seed, scale):
        """SR artifacts over the samples of a generated batch, in place (see ``sample_batch``)."""
        eng = self.engine(tuple(self.shape))
        for b in range(img.shape[0]):
            if sample_ids is not None:
                from ..sharding import sample_seed

                sd32 = sample_seed(base_seed or 0, sample_ids[b])
                np.random.seed(sd32)
                torch.manual_seed(sd32)
            out, meta = self._run_artifacts(img[b], seg[b], genparams)
            if out.data_ptr() != img[b].data_ptr():
                img[b].copy_(out.view(img[b].shape))
            params[b]["artifacts"] = meta
            if scale:
                eng.scale_intensity(img[b], img[b])
        return img, seg, params

    # ------------------------------------------------------------------ batched fast path
    def sample_batch(self, segmentations, seeds, scale: bool = False, out_img=None, out_seg=None, genparams: dict = {}, sample_ids=None, base_seed: int | None = None,
                     artifacts: bool = False):
        """Generate ``len(segmentations)`` independent samples with batched launches.

        segmentations[b]: uint8 device volume; seeds[b]: the reference's seed-path dictionary
        ``{n_sub: {mlabel: path}}`` or a list of 1..4 int8 device volumes (already selected).
        Returns (images [B,*shape] float32, segmentations [B,*shape] uint8, [params]); both
        tensors stay on the device.  ``artifacts=True`` applies the configured SR artifacts (BlurCortex, StructNoise,
        SimulateMotion, SimulatedBoundaries: model.py:200-229) to every sample of the batch after the batched base
        path, on the device, with their own probabilities, and ScaleIntensity (when ``scale``) after them like
        ``FetalSynthDataset.__getitem__`` (datasets.py:311); their draws are reseeded per sample from
        (base_seed, sample id) when ids are given.

        sample_ids + base_seed (multi-GPU streams, ``sharding.py``): every draw of sample b becomes
        a function of (base_seed, sample_ids[b]) — numpy is reseeded per sample and the Philox key
        is (base_seed, sample id) — so the output does not depend on how ids are spread over ranks."""
        if artifacts and any(a is not None for a in self.artifacts.values()):
            img, seg, params = self.sample_batch(segmentations, seeds, False, out_img, out_seg, genparams, sample_ids, base_seed)
            return self._batch_artifacts(img, seg, params, genparams, sample_ids, base_seed, scale)
        shape = tuple(self.shape)
        eng = self.engine(shape)
        plans, params, vols = [], [], []
        if sample_ids is not None and not genparams:
            # throughput path: every sample of the step drawn at once (batch_draw.py), a pure function of
            # (base_seed, sample id)
            from ..batch_draw import draw_batch
            from ..batch_step import run_base_batch, run_step_native
            from .. import _lib as _L
            from ..data.packed import PackedSeeds

            use_dict = any(isinstance(sd, (dict, PackedSeeds)) for sd in seeds)
            d = draw_batch(self, list(sample_ids), int(base_seed or 0), shape, with_subclusters=use_dict)
            B = len(segmentations)
            shp = (B, *shape)
            fast_ok = _FAST_STEP and 1 <= B == len(seeds) and all(t is None or (tuple(t.shape) == shp and t.is_contiguous() and t.device == eng.device) for t in (out_img, out_seg))
            img = seg = None
            if fast_ok:
                img = torch.empty(shp, dtype=torch.float32, device=eng.device) if out_img is None else out_img
                seg = torch.empty(shp, dtype=torch.uint8, device=eng.device) if out_seg is None else out_seg
                fast_ok = img.dtype == torch.float32 and seg.dtype == torch.uint8
            # native step (draws, sample structs, job structs and launches in the library) unless the per-entry-point
            # timing of bench.py is on
            if fast_ok and _NATIVE_STEP and not _L.stats.timing and run_step_native(eng, self, d, seeds, segmentations, img, seg, scale):
                return img, seg, d.params()
            for b, sd in enumerate(seeds):
                if isinstance(sd, (dict, PackedSeeds)):
                    m2s = {m: int(d.m2s[b, m - 1]) for m in range(1, self.intensity_generator.meta_labels + 1)}
                    if isinstance(sd, PackedSeeds):
                        vols.append((sd, m2s))  # labels decoded inside fsg_gmm from the packed words
                    else:
                        vols.append([v.view(-1) for v in self.intensity_generator.select_seeds(sd, m2s, eng.device)])
                else:
                    vols.append([v.view(-1) for v in sd])
            if fast_ok and run_base_batch(eng, d, vols, segmentations, img, seg, scale):
                return img, seg, d.params()
            # anything the vectorised builder does not take: per-sample plans through the generic path
            img, seg = eng.run_base(d.plans(), vols, [s.view(-1) for s in segmentations], out_img=out_img, out_seg=out_seg, scale=scale)
            return img, seg, d.params()
        for b in range(len(segmentations)):
            if sample_ids is not None:
                from ..sharding import sample_seed

                sd32 = sample_seed(base_seed or 0, sample_ids[b])
                np.random.seed(sd32)
                torch.default_generator.manual_seed(sd32)  # CPU generator only: the small tensors (mus, control grids)
                plan = SamplePlan(rng_seed=int(base_seed or 0), sample_id=int(sample_ids[b]))
            else:
                plan = self._new_plan()
            sd = seeds[b]
            if isinstance(sd, dict) or type(sd).__name__ == "PackedSeeds":
                pr = self._draw_generate(plan, sd, shape, genparams, None, device_grids=True)
                vols.append([v.view(-1) for v in plan.meta["seed_vols"]])
            else:
                pr = self._draw_generate(plan, None, shape, genparams, None, device_grids=True)
                plan.mus, plan.sigmas = self.intensity_generator.draw_gmm(genparams.get("seed_intensities", {}))
                vols.append([v.view(-1) for v in sd])
            pr.update(self._draw_augment(plan, shape, genparams, None, device_grids=True))
            plans.append(plan)
            params.append(pr)
        img, seg = eng.run_base(plans, vols, [s.view(-1) for s in segmentations], out_img=out_img, out_seg=out_seg, scale=scale)
        return img, seg, params
