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
from .tables import resample_stds

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


def draw_plans(gen, sample_ids, base_seed: int, shape, with_subclusters: bool = False):
    """Plans + parameter dictionaries for ``len(sample_ids)`` samples of generator ``gen``.
    Control grids are not drawn here: the plans carry (size, std) and the engine draws them on the
    device (``fsg_draw_grids``).  Returns (plans, params[, mlabel2subclusters])."""
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

    # ---- GMM tables (rand_gmm.py:120-145)
    mus = (np.float32(25) + np.float32(200) * U[:, o_mus : o_mus + nlabels].astype(np.float32)).astype(np.float32)
    sigmas = (np.float32(5) + np.float32(20) * U[:, o_sig : o_sig + nlabels].astype(np.float32)).astype(np.float32)
    if tied:
        pert = ndtri(U[:, o_pert : o_pert + nsamp]).astype(np.float32)
        t = mus[:, np.asarray(ig.generation_classes)] + np.float32(25) * pert
        mus[:, np.asarray(ig.seed_labels)] = np.clip(t, np.float32(0), np.float32(225))

    # ---- spatial deformation (affine_nonrigid.py:140-145, 249-324)
    deform_on = S[:, 0] < sd.prob
    flip = S[:, 1] < sd.flip_prb
    rot = (2 * sd.max_rotation * S[:, 2:5] - sd.max_rotation) / 180.0 * np.pi
    shear = 2 * sd.max_shear * S[:, 5:8] - sd.max_shear
    scal = 1 + (2 * sd.max_scaling * S[:, 8:11] - sd.max_scaling)
    A = affine_matrices(rot, shear, scal).astype(np.float32)
    centre2, max_shift, center, shp_f = sd._shape_constants(shape)
    c2 = centre2[None, :] + (2 * (max_shift[None, :] * S[:, 11:14]) - max_shift[None, :])
    nonlin_scale = sd.nonlin_scale_min + S[:, 14] * (sd.nonlin_scale_max - sd.nonlin_scale_min)
    size_f = np.round(nonlin_scale[:, None] * shp_f[None, :]).astype(int)
    nonlin_std = sd.nonlin_std_max * S[:, 15]

    # ---- gamma (synthseg.py:262-275), bias field (:157-176), resolution (:63-80), noise (:217-235)
    gamma_on = S[:, 16] < gm.prob
    gamma = np.exp(gm.gamma_std * ndtri(S[:, 17]))
    bias_on = S[:, 18] < bf.prob
    bf_scale = bf.scale_min + S[:, 19] * (bf.scale_max - bf.scale_min)
    bf_size = np.maximum(np.round(bf_scale[:, None] * np.asarray(shape)[None, :]).astype(int), 1)
    bf_std = bf.std_min + (bf.std_max - bf.std_min) * S[:, 20]
    res_on = S[:, 21] < rs_.prob
    spacing = rs_.min_resolution + (rs_.max_resolution - rs_.min_resolution) * S[:, 22]
    blur_u = S[:, 23]
    noise_on = S[:, 24] < nz.prob
    noise_std = nz.std_min + (nz.std_max - nz.std_min) * S[:, 25]
    m2s = None
    if with_subclusters:  # rand_gmm.py:81-85: randint(min, max + 1) per meta label
        k = ig.max_subclusters - ig.min_subclusters + 1
        m2s = ig.min_subclusters + np.minimum((S[:, 26 : 26 + ig.meta_labels] * k).astype(int), k - 1)

    res = np.asarray(gen.resolution, dtype=np.float64)
    plans, params = [], []
    for b in range(B):
        p = SamplePlan(mus=mus[b], sigmas=sigmas[b], rng_seed=int(base_seed), sample_id=int(sample_ids[b]))
        pr = {"selected_seeds": {}, "seed_intensities": {}}
        if deform_on[b]:
            p.deform, p.flip, p.A, p.c2, p.center = True, bool(flip[b]), A[b], c2[b].astype(np.float64), center
            non_rigid = {}
            if sd.nonlinear_transform:
                p.fsmall_dev = (tuple(int(v) for v in size_f[b]), float(np.float32(nonlin_std[b])))
                non_rigid = {"nonlin_scale": np.array([nonlin_scale[b]]), "nonlin_std": float(nonlin_std[b]), "size_F_small": size_f[b].tolist()}
            pr["deform_params"] = {"affine": {"rotations": rot[b], "shears": shear[b], "scalings": scal[b]}, "non_rigid": non_rigid, "flip": bool(flip[b])}
        else:
            pr["deform_params"] = {"affine": None, "non_rigid": None, "flip": False}
        if gamma_on[b]:
            p.gamma = float(gamma[b])
        pr["gamma_params"] = {"gamma": p.gamma}
        if bias_on[b]:
            p.bf_dev = (tuple(int(v) for v in bf_size[b]), float(np.float32(bf_std[b])))
            pr["bf_params"] = {"bf_scale": np.array([bf_scale[b]]), "bf_std": np.array([bf_std[b]]), "bf_size": bf_size[b].tolist()}
        else:
            pr["bf_params"] = {"bf_scale": None, "bf_std": None, "bf_size": None}
        if res_on[b]:
            p.spacing = np.array([spacing[b]] * 3, dtype=np.float64)
            p.stds = resample_stds(p.spacing, res, float(blur_u[b]))
        pr["resample_params"] = {"spacing": None if p.spacing is None else p.spacing.tolist()}
        if noise_on[b]:
            p.noise_std = float(np.float32(noise_std[b]))
        pr["noise_params"] = {"noise_std": p.noise_std}
        plans.append(p)
        params.append(pr)
    return (plans, params, m2s) if with_subclusters else (plans, params)
