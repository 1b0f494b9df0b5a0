"""Vectorised parameter draws for the batched generator.

The stage classes (``generator/**``) draw one sample's parameters with ~40 small numpy / torch
calls in the reference's RNG order — needed where the reference's scalar stream must be reproduced
(``FetalSynthGen.sample / generate / augment``), but ~150 us of Python per sample.  The batched
throughput path (``FetalSynthGen.sample_batch`` with sample ids) has no reference stream to follow:
its contract is the *distributions* and the independence from the sharding, so all samples of a
step are drawn at once from a counter-based generator — uniform ``(sample, column)`` =
splitmix64(base_seed, sample_id, column) — and shaped with a few array operations.  Every
distribution below cites the stage code it mirrors; ``tests/test_host.py`` compares the two
statistically.
"""
from __future__ import annotations

import numpy as np
from scipy.special import ndtri

from .engine import SamplePlan

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
    return x ^ (x >> np.uint64(31))


def uniforms(base_seed: int, sample_ids, ncols: int) -> np.ndarray:
    """[len(sample_ids), ncols] float64 uniforms in (0, 1), a pure function of (base_seed, id, column)."""
    with np.errstate(over="ignore"):
        ids = np.asarray(sample_ids, dtype=np.uint64)[:, None]
        cols = np.arange(ncols, dtype=np.uint64)[None, :]
        key = _splitmix(np.uint64(base_seed & 0xFFFFFFFFFFFFFFFF) ^ _splitmix(ids))
        bits = _splitmix(key + cols * np.uint64(0xD1342543DE82EF95))
    return ((bits >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def affine_matrices(rot: np.ndarray, sh: np.ndarray, sc: np.ndarray) -> np.ndarray:
    """Batched ``make_affine_matrix`` (utils/generation.py:39-71): [B,3,3] float64."""
    B = rot.shape[0]
    c, s = np.cos(rot), np.sin(rot)
    m = np.zeros((6, B, 3, 3), dtype=np.float64)
    m[:, :, 0, 0] = m[:, :, 1, 1] = m[:, :, 2, 2] = 1.0
    m[0, :, 1, 0], m[0, :, 2, 0] = sh[:, 1], sh[:, 2]
    m[1, :, 0, 1], m[1, :, 2, 1] = sh[:, 0], sh[:, 2]
    m[2, :, 0, 2], m[2, :, 1, 2] = sh[:, 0], sh[:, 1]
    m[3, :, 1, 1], m[3, :, 1, 2], m[3, :, 2, 1], m[3, :, 2, 2] = c[:, 0], -s[:, 0], s[:, 0], c[:, 0]
    m[4, :, 0, 0], m[4, :, 0, 2], m[4, :, 2, 0], m[4, :, 2, 2] = c[:, 1], s[:, 1], -s[:, 1], c[:, 1]
    m[5, :, 0, 0], m[5, :, 0, 1], m[5, :, 1, 0], m[5, :, 1, 1] = c[:, 2], -s[:, 2], s[:, 2], c[:, 2]
    a = m[0]
    for k in range(1, 6):
        a = a @ m[k]
    return a * sc[:, :, None]


class BatchDraw:
    """Everything ``draw_plans`` draws for the samples of one step, as arrays over the batch (the vectorised
    launch builder ``batch_step.run_base_batch`` consumes them directly; ``plans()`` / ``params()`` give the
    per-sample ``SamplePlan`` objects and the reference-shaped parameter dictionaries)."""

    __slots__ = ("B", "base_seed", "sample_ids", "shape", "res", "mus", "sigmas", "deform_on", "flip", "nonlinear", "rot", "shear", "scal", "A", "c2",
                 "center", "nonlin_scale", "size_f", "nonlin_std", "gamma_on", "gamma", "bias_on", "bf_scale", "bf_size", "bf_std", "res_on", "spacing",
                 "stds", "noise_on", "noise_std", "m2s", "_c_out")

    def plans(self):
        out = []
        for b in range(self.B):
            p = SamplePlan(mus=self.mus[b], sigmas=self.sigmas[b], rng_seed=int(self.base_seed), sample_id=int(self.sample_ids[b]))
            if self.deform_on[b]:
                p.deform, p.flip, p.A, p.c2, p.center = True, bool(self.flip[b]), self.A[b], self.c2[b].astype(np.float64), self.center
                if self.nonlinear:
                    p.fsmall_dev = (tuple(int(v) for v in self.size_f[b]), float(self.nonlin_std[b]))
            if self.gamma_on[b]:
                p.gamma = float(self.gamma[b])
            if self.bias_on[b]:
                p.bf_dev = (tuple(int(v) for v in self.bf_size[b]), float(self.bf_std[b]))
            if self.res_on[b]:
                p.spacing = np.array([self.spacing[b]] * 3, dtype=np.float64)
                p.stds = self.stds[b].copy()
            if self.noise_on[b]:
                p.noise_std = float(self.noise_std[b])
            out.append(p)
        return out

    def param(self, b: int) -> dict:
        """Parameter dictionary of sample b in the shape ``FetalSynthGen.sample`` returns (model.py:231-257)."""
        pr = {"selected_seeds": {}, "seed_intensities": {}}
        if self.deform_on[b]:
            non_rigid = {}
            if self.nonlinear:
                non_rigid = {"nonlin_scale": np.array([self.nonlin_scale[b]]), "nonlin_std": float(self.nonlin_std[b]),
                             "size_F_small": self.size_f[b].tolist()}
            pr["deform_params"] = {"affine": {"rotations": self.rot[b], "shears": self.shear[b], "scalings": self.scal[b]}, "non_rigid": non_rigid, "flip": bool(self.flip[b])}
        else:
            pr["deform_params"] = {"affine": None, "non_rigid": None, "flip": False}
        pr["gamma_params"] = {"gamma": float(self.gamma[b]) if self.gamma_on[b] else None}
        if self.bias_on[b]:
            pr["bf_params"] = {"bf_scale": np.array([self.bf_scale[b]]), "bf_std": np.array([self.bf_std[b]]), "bf_size": self.bf_size[b].tolist()}
        else:
            pr["bf_params"] = {"bf_scale": None, "bf_std": None, "bf_size": None}
        pr["resample_params"] = {"spacing": [float(self.spacing[b])] * 3 if self.res_on[b] else None}
        pr["noise_params"] = {"noise_std": float(self.noise_std[b]) if self.noise_on[b] else None}
        if self.m2s is not None:
            pr["selected_seeds"] = {"mlabel2subclusters": {m + 1: int(v) for m, v in enumerate(self.m2s[b])}}
        return pr

    def params(self):
        return LazyParams(self)


class LazyParams:
    """List-like view of the per-sample parameter dictionaries, built on access (a training loop that never looks
    at them does not pay for eight dictionaries per step)."""

    def __init__(self, draw: BatchDraw):
        self._d, self._cache = draw, {}

    def __len__(self):
        return self._d.B

    def __getitem__(self, b):
        if isinstance(b, slice):
            return [self[i] for i in range(*b.indices(len(self)))]
        if b < 0:
            b += self._d.B
        if not 0 <= b < self._d.B:
            raise IndexError(b)
        if b not in self._cache:
            self._cache[b] = self._d.param(b)
        return self._cache[b]

    def __iter__(self):
        return (self[b] for b in range(self._d.B))


def _draw_config(gen, shape):
    """``fsg_draw_config`` of a generator for volumes of `shape` (cached on the generator)."""
    import ctypes as C

    from . import _lib

    cache = gen.__dict__.setdefault("_draw_cfg", {})
    hit = cache.get(shape)
    if hit is not None:
        return hit
    ig, sd, rs_, bf, nz, gm = gen.intensity_generator, gen.spatial_deform, gen.resampled, gen.biasfield, gen.noise, gen.gamma
    c = _lib.DrawConfig()
    labels = np.ascontiguousarray(ig.seed_labels, dtype=np.int32)
    classes = np.ascontiguousarray(ig.generation_classes, dtype=np.int32)
    c.nlabels, c.nseed = int(max(ig.seed_labels)) + 1, len(ig.seed_labels)
    c.seed_labels, c.generation_classes = labels.ctypes.data, classes.ctypes.data
    c.tied = int(ig.generation_classes != ig.seed_labels)
    c.meta_labels, c.min_subclusters, c.max_subclusters = int(ig.meta_labels), int(ig.min_subclusters), int(ig.max_subclusters)
    c.shape = (C.c_int32 * 3)(*shape)
    c.nonlinear = int(bool(sd.nonlinear_transform))
    c.res = (C.c_double * 3)(*[float(v) for v in gen.resolution])
    c.deform_prob, c.flip_prb, c.max_rotation, c.max_shear, c.max_scaling = float(sd.prob), float(sd.flip_prb), float(sd.max_rotation), float(sd.max_shear), float(sd.max_scaling)
    c.nonlin_scale_min, c.nonlin_scale_max, c.nonlin_std_max = float(sd.nonlin_scale_min), float(sd.nonlin_scale_max), float(sd.nonlin_std_max)
    centre2, max_shift, center, _ = sd._shape_constants(shape)
    c.centre2 = (C.c_double * 3)(*[float(v) for v in centre2])
    c.max_shift = (C.c_double * 3)(*[float(v) for v in max_shift])
    c.gamma_prob, c.gamma_std = float(gm.prob), float(gm.gamma_std)
    c.bias_prob, c.bf_scale_min, c.bf_scale_max, c.bf_std_min, c.bf_std_max = float(bf.prob), float(bf.scale_min), float(bf.scale_max), float(bf.std_min), float(bf.std_max)
    c.res_prob, c.min_resolution, c.max_resolution = float(rs_.prob), float(rs_.min_resolution), float(rs_.max_resolution)
    c.noise_prob, c.noise_std_min, c.noise_std_max = float(nz.prob), float(nz.std_min), float(nz.std_max)
    hit = cache[shape] = (c, center, (labels, classes))
    return hit


_OUT_FIELDS = (("mus", np.float32, "L"), ("sigmas", np.float32, "L"), ("deform_on", np.bool_, 0), ("flip", np.bool_, 0), ("gamma_on", np.bool_, 0), ("bias_on", np.bool_, 0),
               ("res_on", np.bool_, 0), ("noise_on", np.bool_, 0), ("rot", np.float64, 3), ("shear", np.float64, 3), ("scal", np.float64, 3), ("A", np.float32, 9),
               ("c2", np.float64, 3), ("nonlin_scale", np.float64, 0), ("size_f", np.int64, 3), ("nonlin_std", np.float32, 0), ("gamma", np.float64, 0),
               ("bf_scale", np.float64, 0), ("bf_size", np.int64, 3), ("bf_std", np.float32, 0), ("spacing", np.float64, 0), ("stds", np.float64, 3),
               ("noise_std", np.float32, 0), ("m2s", np.int64, "M"))


def draw_batch(gen, sample_ids, base_seed: int, shape, with_subclusters: bool = False) -> BatchDraw:
    """All parameters of ``len(sample_ids)`` samples of generator ``gen``, a pure function of (base_seed, id), drawn
    by the library (``fsg_draw_batch``, csrc/step.cu: the one definition shared by the native step and the per-sample
    plans).  Control grids are not drawn here: the draw carries (size, std) and the engine draws them on the device
    (``fsg_draw_grids``)."""
    from . import _lib

    shape = tuple(int(v) for v in shape)
    cfg, center, _ = _draw_config(gen, shape)
    B = len(sample_ids)
    d = BatchDraw()
    d.B, d.base_seed, d.sample_ids, d.shape = B, int(base_seed), np.ascontiguousarray(sample_ids, dtype=np.uint64), shape
    d.center, d.nonlinear, d.res = center, bool(cfg.nonlinear), np.asarray(gen.resolution, dtype=np.float64)
    out = _lib.DrawOut()
    for name, dt, k in _OUT_FIELDS:
        if name == "m2s" and not with_subclusters:
            d.m2s = None
            continue
        k = cfg.nlabels if k == "L" else (cfg.meta_labels if k == "M" else k)
        a = np.empty((B, k) if k else (B,), dtype=dt)
        setattr(d, name, a)
        setattr(out, name, a.ctypes.data)
    import ctypes as C

    rc = _lib.load().fsg_draw_batch(C.byref(cfg), d.sample_ids.ctypes.data, B, int(base_seed) & (2**64 - 1), C.byref(out))
    if rc:
        raise _lib.FsgError(f"fsg_draw_batch failed ({rc}): {_lib.load().fsg_last_error().decode()}")
    d.A = d.A.reshape(B, 3, 3)
    d._c_out = out  # fsg_draw_out over the arrays above (batch_step.run_step_native hands it to fsg_step_fill)
    return d


def draw_batch_numpy(gen, sample_ids, base_seed: int, shape, with_subclusters: bool = False) -> BatchDraw:
    """The same draws in numpy (the definition ``fsg_draw_batch`` was written from; kept for the test that compares
    the two: equal booleans and integers, floats to rounding of cos / sin / exp / ndtri)."""
    ig, sd, rs_, bf, nz, gm = gen.intensity_generator, gen.spatial_deform, gen.resampled, gen.biasfield, gen.noise, gen.gamma
    B = len(sample_ids)
    shape = tuple(int(v) for v in shape)
    nlabels = max(ig.seed_labels) + 1
    nsamp = len(ig.seed_labels)
    tied = ig.generation_classes != ig.seed_labels
    o_mus, o_sig, o_pert = 0, nlabels, 2 * nlabels
    o_s = 2 * nlabels + nsamp  # scalar block
    U = uniforms(base_seed, sample_ids, o_s + 40)
    S = U[:, o_s:]
    d = BatchDraw()
    d.B, d.base_seed, d.sample_ids, d.shape = B, int(base_seed), np.asarray(sample_ids, dtype=np.uint64), shape

    # ---- GMM tables (rand_gmm.py:120-145)
    mus = (np.float32(25) + np.float32(200) * U[:, o_mus : o_mus + nlabels].astype(np.float32)).astype(np.float32)
    sigmas = (np.float32(5) + np.float32(20) * U[:, o_sig : o_sig + nlabels].astype(np.float32)).astype(np.float32)
    if tied:
        pert = ndtri(U[:, o_pert : o_pert + nsamp]).astype(np.float32)
        t = mus[:, np.asarray(ig.generation_classes)] + np.float32(25) * pert
        mus[:, np.asarray(ig.seed_labels)] = np.clip(t, np.float32(0), np.float32(225))
    d.mus, d.sigmas = mus, sigmas

    # ---- spatial deformation (affine_nonrigid.py:140-145, 249-324)
    d.deform_on = S[:, 0] < sd.prob
    d.flip = S[:, 1] < sd.flip_prb
    d.rot = (2 * sd.max_rotation * S[:, 2:5] - sd.max_rotation) / 180.0 * np.pi
    d.shear = 2 * sd.max_shear * S[:, 5:8] - sd.max_shear
    d.scal = 1 + (2 * sd.max_scaling * S[:, 8:11] - sd.max_scaling)
    d.A = affine_matrices(d.rot, d.shear, d.scal).astype(np.float32)
    centre2, max_shift, center, shp_f = sd._shape_constants(shape)
    d.center = center
    d.c2 = centre2[None, :] + (2 * (max_shift[None, :] * S[:, 11:14]) - max_shift[None, :])
    d.nonlinear = bool(sd.nonlinear_transform)
    d.nonlin_scale = sd.nonlin_scale_min + S[:, 14] * (sd.nonlin_scale_max - sd.nonlin_scale_min)
    d.size_f = np.round(d.nonlin_scale[:, None] * shp_f[None, :]).astype(int)
    d.nonlin_std = (sd.nonlin_std_max * S[:, 15]).astype(np.float32)  # the kernels take float32 scales

    # ---- gamma (synthseg.py:262-275), bias field (:157-176), resolution (:63-80), noise (:217-235)
    d.gamma_on = S[:, 16] < gm.prob
    d.gamma = np.exp(gm.gamma_std * ndtri(S[:, 17]))
    d.bias_on = S[:, 18] < bf.prob
    d.bf_scale = bf.scale_min + S[:, 19] * (bf.scale_max - bf.scale_min)
    d.bf_size = np.maximum(np.round(d.bf_scale[:, None] * np.asarray(shape)[None, :]).astype(int), 1)
    d.bf_std = (bf.std_min + (bf.std_max - bf.std_min) * S[:, 20]).astype(np.float32)
    d.res_on = S[:, 21] < rs_.prob
    d.spacing = rs_.min_resolution + (rs_.max_resolution - rs_.min_resolution) * S[:, 22]
    d.res = np.asarray(gen.resolution, dtype=np.float64)
    # blur widths of the resolution simulation (tables.resample_stds, synthseg.py:78-80), all samples at once
    sp3 = np.repeat(d.spacing[:, None], 3, axis=1)
    d.stds = (0.85 + 0.3 * S[:, 23:24]) * np.log(5) / np.pi * sp3 / d.res[None, :]
    d.stds[sp3 <= d.res[None, :]] = 0.0
    d.noise_on = S[:, 24] < nz.prob
    d.noise_std = (nz.std_min + (nz.std_max - nz.std_min) * S[:, 25]).astype(np.float32)
    d.m2s = None
    if with_subclusters:  # rand_gmm.py:81-85: randint(min, max + 1) per meta label
        k = ig.max_subclusters - ig.min_subclusters + 1
        d.m2s = ig.min_subclusters + np.minimum((S[:, 26 : 26 + ig.meta_labels] * k).astype(int), k - 1)
    return d


def draw_plans(gen, sample_ids, base_seed: int, shape, with_subclusters: bool = False):
    """Plans + parameter dictionaries for ``len(sample_ids)`` samples of generator ``gen``.
    Returns (plans, params[, mlabel2subclusters])."""
    d = draw_batch(gen, sample_ids, base_seed, shape, with_subclusters)
    plans, params = d.plans(), [d.param(b) for b in range(d.B)]
    return (plans, params, d.m2s) if with_subclusters else (plans, params)
