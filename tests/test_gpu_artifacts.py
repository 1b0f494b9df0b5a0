"""GPU parity tests of the SR-artifact kernels (BlurCortex, StructNoise, SimulatedBoundaries)
against golden vectors of the unmodified reference classes and the numpy oracle."""
import numpy as np
import pytest
import torch

import np_artifacts as OA
from golden_util import GOLDEN, load_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4

if torch.cuda.is_available():
    from fetalsyngen_b200 import _lib
    from fetalsyngen_b200.artifact_ops import ArtifactOps
    from fetalsyngen_b200.engine import engine_for
    from fetalsyngen_b200.generator.artifacts.utils import StructNoiseMergeParams
    from fetalsyngen_b200.generator.augmentation.artifacts import BlurCortex, SimulatedBoundaries, StructNoise


def art(name):
    with np.load(GOLDEN / f"art_{name}.npz", allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def base():
    d = load_case("c64_default")
    return d["final"].astype(np.float32), d["seg_out"]


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    rng = float(b.max() - b.min()) or 1.0
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) / rng


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def test_mog_kernel_vs_reference():
    g = art("blur_cortex")
    img, _ = base()
    eng = engine_for(DEV, img.shape, (1.0, 1.0, 1.0))
    ops = ArtifactOps(eng)
    c = dev(g["centers"].astype(np.float32)[:, ::-1].copy())
    s = dev(g["sigmas"].astype(np.float32)[:, ::-1].copy())
    out = torch.empty(eng.nvox, dtype=torch.float32, device=DEV)
    ops.mog(c, s, out=out)
    assert np.abs(out.cpu().numpy().reshape(img.shape) - g["gaussian"]).max() <= 2e-5


def test_blur_cortex_vs_reference():
    g = art("blur_cortex")
    img, seg = base()
    bc = BlurCortex(prob=1.0, cortex_label=2, nblur_min=50, nblur_max=200)
    out, meta = bc(dev(img), dev(seg), DEV, {}, inject={"nblur": g["nblur"], "std_blurs": g["std_blurs"], "centers": g["centers"], "sigmas": g["sigmas"]})
    assert meta["nblur"] == int(g["nblur"])
    assert rel(out, g["output"]) <= TOL


def test_blur_cortex_philox_mode_blurs_only_near_cortex():
    img, seg = base()
    np.random.seed(0)
    torch.manual_seed(0)
    bc = BlurCortex(prob=1.0, cortex_label=2, nblur_min=50, nblur_max=200)
    out, meta = bc(dev(img), dev(seg), DEV)
    out = out.cpu().numpy()
    assert np.isfinite(out).all() and 50 <= meta["nblur"] < 200
    changed = np.abs(out - img) > 1e-4
    assert changed.any()
    # blobs are centred on (transposed) cortex voxels with sigma ~ Gamma(3,1): the change is local
    assert changed.mean() < 0.9
    assert out.min() >= -1e-6 and out.max() <= img.max() + 1e-5


@pytest.mark.parametrize("name", ["struct_noise", "struct_noise_oct"])
def test_struct_noise_vs_reference(name):
    g = art(name)
    img, seg = base()
    mp = StructNoiseMergeParams(merge_type="perlin", perlin_res_list=[1, 2], perlin_octaves_list=[1, 2, 4], perlin_persistence=0.5, perlin_lacunarity=2, perlin_increase_size=0.1)
    sn = StructNoise(prob=1.0, wm_label=3, std_min=0.2, std_max=0.4, merge_params=mp, nstages_min=1, nstages_max=5)
    inj = {k: g[k] for k in g if k.startswith(("randn_", "theta_", "phi_"))}
    inj.update(nstages=g["nstages"], noise_std=g["noise_std"], res=g["res"], octave=g["octave"])
    out, meta = sn(dev(img), dev(seg), DEV, {}, inject=inj)
    assert meta["nstages"] == int(g["nstages"]) and meta["octave"] == int(g["octave"])
    assert rel(out, g["output"]) <= TOL
    # pieces: raw fractal noise -> normalised weight, noise pyramid
    eng = engine_for(DEV, img.shape, (1.0, 1.0, 1.0))
    ops = ArtifactOps(eng)
    w = torch.empty(eng.nvox, dtype=torch.float32, device=DEV)
    mm = torch.zeros(2, dtype=torch.float32, device=DEV)
    StructNoise.perlin_weight(eng, ops, int(g["res"]), int(g["octave"]), 0.5, 2, w, mm, inj)
    lo, hi = mm.cpu().numpy()
    weight = np.clip((w.cpu().numpy().reshape(img.shape) + 0.1 - lo) / (hi - lo), 0, 1)
    assert np.abs(weight - g["weight"]).max() <= 1e-4
    lr = torch.empty(eng.nvox, dtype=torch.float32, device=DEV)
    sn.multiscale_noise(eng, ops, int(g["nstages"]), lr, mm, inj)
    want = OA.multiscale_noise(img.shape, [g[f"randn_{k}"] for k in range(int(g["nstages"]))])
    lo, hi = mm.cpu().numpy()
    got = lr.cpu().numpy().reshape(img.shape) / max(abs(lo), abs(hi))
    assert np.abs(got - want).max() <= 1e-5


def test_struct_noise_philox_and_gaussian_merge():
    img, seg = base()
    np.random.seed(1)
    torch.manual_seed(1)
    for merge in ("perlin", "gaussian"):
        mp = StructNoiseMergeParams(merge_type=merge, gauss_nloc_min=5, gauss_nloc_max=15, gauss_sigma_mu=25, gauss_sigma_std=5, perlin_res_list=[1, 2],
                                    perlin_octaves_list=[1, 2, 4], perlin_persistence=0.5, perlin_lacunarity=2, perlin_increase_size=0.1)
        sn = StructNoise(prob=1.0, wm_label=3, std_min=0.2, std_max=0.4, merge_params=mp)
        out, meta = sn(dev(img), dev(seg), DEV)
        out = out.cpu().numpy()
        assert np.isfinite(out).all()
        assert np.array_equal(out[seg == 0], img[seg == 0])          # only inside the brain
        assert (np.abs(out - img)[seg > 0] > 1e-6).mean() > 0.2
        assert out.min() >= 0 and out.max() <= 2 * img.max() + 1e-5


def _keep_masks(g, img, seg):
    """Keep-masks of the fuzzy rounds from the reference's randperm draws."""
    keep = {}
    m = g["mask_halo"]
    for i in range(int(g["n_generate_fuzzy"])):
        diff = OA.dilate(m, 7).astype(np.int32) - m.astype(np.int32)
        nz = np.nonzero(diff)
        n = len(nz[0])
        drop = g[f"perm_{i}"][: int(n * 0.9)]
        k = diff.astype(np.uint8)
        k[nz[0][drop], nz[1][drop], nz[2][drop]] = 0
        keep[f"keep_{i}"] = k
        m = g[f"fuzzy_{i}"]
    return keep


def test_boundaries_morphology_bit_exact():
    g = art("boundaries")
    img, seg = base()
    eng = engine_for(DEV, img.shape, (1.0, 1.0, 1.0))
    ops = ArtifactOps(eng)
    sb = SimulatedBoundaries(0.0, 1.0, 1.0)
    mask = dev((seg > 0).astype(np.uint8)).view(-1)
    halo = torch.empty_like(mask)
    sb.build_halo(ops, mask, int(g["halo_radius"]), halo)
    np.testing.assert_array_equal(halo.cpu().numpy().reshape(img.shape), g["mask_halo"])
    keep = _keep_masks(g, img, seg)
    cur = halo
    for i in range(int(g["n_generate_fuzzy"])):
        nxt = torch.empty_like(mask)
        sb.generate_fuzzy_boundaries(ops, cur, nxt, dev(keep[f"keep_{i}"]).view(-1))
        np.testing.assert_array_equal(nxt.cpu().numpy().reshape(img.shape), g[f"fuzzy_{i}"])
        cur = nxt
    # box ops against the oracle on a random mask (zero-padded erosion at the border included)
    rs = np.random.RandomState(0)
    m = (rs.rand(*img.shape) > 0.7).astype(np.uint8)
    a, b, c = dev(m).view(-1), torch.empty_like(mask), torch.empty_like(mask)
    for k in (3, 5, 7):
        np.testing.assert_array_equal(ops.box(a, b, c, k, 0).cpu().numpy().reshape(m.shape), OA.dilate(m, k))
        np.testing.assert_array_equal(ops.box(a, b, c, k, 1).cpu().numpy().reshape(m.shape), OA.erode(m, k))
    # L1 distance thresholds == repeated 6-neighbour dilations
    sparse = (rs.rand(*img.shape) > 0.995).astype(np.uint8)
    d16, t16 = ops.u16("d0"), ops.u16("d1")
    ops.dist(dev(sparse).view(-1), d16, t16, 5, 1)
    d = d16.cpu().numpy().view(np.uint16).reshape(img.shape)
    want = sparse
    for kk in range(1, 5):
        want = OA.build_halo(want, 1)
        np.testing.assert_array_equal((d <= kk).astype(np.uint8), want)


def test_boundaries_vs_reference():
    g = art("boundaries")
    img, seg = base()
    inj = {k: g[k] for k in ("halo_radius", "n_generate_fuzzy", "n_centers", "base_sigma", "centers", "sigmas")}
    inj.update(_keep_masks(g, img, seg))
    sb = SimulatedBoundaries(0.0, 1.0, 1.0)
    out, meta = sb(dev(img), dev(seg), DEV, {}, inject=inj)
    out = out.cpu().numpy()
    assert meta == {"no_mask_on": False, "halo_on": True, "fuzzy_on": True}
    # the mask depends on round(p * len - 1) of a float MoG: allow flips only where p*len-1 sits on a rounding boundary
    mism = out != g["output"]
    frac = g["mog"] * (6 * (int(g["n_generate_fuzzy"]) - 1)) - 1
    near_tie = np.abs(frac - np.floor(frac) - 0.5) < 1e-3
    assert not (mism & ~near_tie).any()
    assert mism.mean() < 1e-4


def test_boundaries_philox_mode_and_gates():
    img, seg = base()
    np.random.seed(4)
    torch.manual_seed(4)
    brain = seg > 0
    for probs in ((0.0, 1.0, 1.0), (0.0, 1.0, 0.0), (0.0, 0.0, 0.0), (1.0, 0.0, 0.0)):
        sb = SimulatedBoundaries(*probs)
        out, meta = sb(dev(img), dev(seg), DEV)
        out = out.cpu().numpy()
        if probs[0] == 1.0:
            assert np.array_equal(out, img) and meta["no_mask_on"]
            continue
        kept = out != 0
        assert np.array_equal(out[kept], img[kept])
        assert (kept | (img == 0))[brain].all()           # the brain itself is never masked out
        if probs[1] == 0.0 and probs[2] == 0.0:
            assert not kept[~brain].any()                   # plain brain mask


def test_sample_voxels_distribution():
    shape = (64, 64, 64)
    eng = engine_for(DEV, shape, (1.0, 1.0, 1.0))
    ops = ArtifactOps(eng)
    lab = np.zeros(shape, dtype=np.uint8)
    lab[8:24, 8:24, 8:24] = 2       # 4096 candidates
    lab[40:44, 40:44, 40:44] = 3
    labd = dev(lab).view(-1)
    hits = np.zeros(shape)
    for trial in range(64):
        c, n = ops.sample_voxels(labd, 100, match=2, transpose_out=False, rng=(99, trial))
        c = c.cpu().numpy().astype(int)
        assert n == 100 and len({tuple(r) for r in c}) == 100      # without replacement
        assert (lab[c[:, 0], c[:, 1], c[:, 2]] == 2).all()
        hits[c[:, 0], c[:, 1], c[:, 2]] += 1
    # uniform: every octant of the cube gets ~1/8 of the 6400 draws
    oct_counts = [hits[8 + 8 * a : 16 + 8 * a, 8 + 8 * b : 16 + 8 * b, 8 + 8 * c : 16 + 8 * c].sum() for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    assert max(abs(o - 800) for o in oct_counts) < 5 * np.sqrt(800)
    # weighted: a prior centred on one corner concentrates the draws there
    prior = ([(8, 8, 8)], [[4, 4, 4]])
    c, n = ops.sample_voxels(labd, 50, match=2, prior=prior, transpose_out=False, rng=(7, 0))
    c = c.cpu().numpy()
    assert n == 50 and np.linalg.norm(c - 8, axis=1).mean() < 8
    # fewer candidates than requested: everything is taken at most once
    c, n = ops.sample_voxels(labd, 100, match=3, transpose_out=False, rng=(7, 1))
    assert n <= 64 and n >= 40


@pytest.mark.parametrize("shape", [(32, 48, 64), (20, 24, 30), (16, 16, 16), (9, 33, 20)])
def test_box_morphology_packed_and_scalar_paths_match_scipy(shape):
    """fsg_morph_box (max / zero-padded min / count) on shapes that take the packed-voxel kernels
    (z extent a multiple of 16) and on shapes that fall back to the voxel-per-thread kernel, including
    half-widths above the packed z kernel's range — all bit-exact against scipy.ndimage."""
    from scipy import ndimage

    eng = engine_for(DEV, shape, (1.0, 1.0, 1.0))
    ops = ArtifactOps(eng)
    rs = np.random.RandomState(shape[0])
    m = (rs.rand(*shape) > 0.6).astype(np.uint8)
    a = dev(m).view(-1)
    b, c = torch.empty_like(a), torch.empty_like(a)
    for k in (3, 5, 7, 9):
        got = ops.box(a, b, c, k, 0).cpu().numpy().reshape(shape)
        np.testing.assert_array_equal(got, ndimage.maximum_filter(m, size=k, mode="constant", cval=0))
        got = ops.box(a, b, c, k, 1).cpu().numpy().reshape(shape)
        np.testing.assert_array_equal(got, ndimage.minimum_filter(m, size=k, mode="constant", cval=0))
    for k in (3, 5):
        got = ops.box(a, b, c, k, 2).cpu().numpy().reshape(shape)
        want = ndimage.uniform_filter(m.astype(np.float64), size=k, mode="constant", cval=0) * k**3
        np.testing.assert_array_equal(got, np.round(want).astype(np.uint8))


@pytest.mark.parametrize("shape", [(32, 48, 64), (20, 24, 30), (8, 8, 8)])
def test_windowed_distance_packed_and_scalar_paths(shape):
    """fsg_morph_dist (squared Euclidean / L1 distance to the nearest set voxel inside a +-r window,
    65535 = none) against a brute-force numpy evaluation, on shapes that take the packed halfword
    kernels for the strided passes and on shapes that do not."""
    eng = engine_for(DEV, shape, (1.0, 1.0, 1.0))
    ops = ArtifactOps(eng)
    rs = np.random.RandomState(shape[1])
    m = (rs.rand(*shape) > 0.97).astype(np.uint8)
    d16, t16 = ops.u16("d0"), ops.u16("d1")
    for r, metric in ((3, 0), (2, 1), (5, 0)):
        want = np.full(shape, 65535, dtype=np.int64)
        pad = np.pad(m, r)
        for dx in range(-r, r + 1):
            for dy in range(-r, r + 1):
                for dz in range(-r, r + 1):
                    c = dx * dx + dy * dy + dz * dz if metric == 0 else abs(dx) + abs(dy) + abs(dz)
                    sh = pad[r + dx : r + dx + shape[0], r + dy : r + dy + shape[1], r + dz : r + dz + shape[2]]
                    want = np.where(sh > 0, np.minimum(want, c), want)
        ops.dist(dev(m).view(-1), d16, t16, r, metric)
        got = d16.cpu().numpy().view(np.uint16).reshape(-1)[: m.size].reshape(shape).astype(np.int64)
        np.testing.assert_array_equal(got, want)


def test_boundaries_morphology_bit_exact_at_256():
    """SimulatedBoundaries' integer morphology at BASELINE.json's full size on the bundled sub-sta30 label map:
    halo (ball dilation == exact integer distance threshold, checked against scipy's Euclidean feature transform),
    one fuzzy round against the oracle, and the L1 dilation stack — all bit-exact."""
    from scipy import ndimage

    from golden_util import load_subject

    seg = load_subject("sub-sta30")[0]
    shape = seg.shape
    m = (seg > 0).astype(np.uint8)
    eng = engine_for(DEV, shape, (0.5, 0.5, 0.5))
    ops = ArtifactOps(eng)
    sb = SimulatedBoundaries(0.0, 1.0, 1.0)
    mask = dev(m).view(-1)
    halo = torch.empty_like(mask)
    r = 9
    sb.build_halo(ops, mask, r, halo)
    idx = ndimage.distance_transform_edt(m == 0, return_distances=False, return_indices=True)
    grid = np.indices(shape)
    d2 = sum((idx[a].astype(np.int64) - grid[a]) ** 2 for a in range(3))
    halo_np = halo.cpu().numpy().reshape(shape)
    np.testing.assert_array_equal(halo_np, (d2 <= r * r).astype(np.uint8))
    # one fuzzy round with an injected permutation of the ring voxels
    diff = OA.dilate(halo_np, 7).astype(np.int32) - halo_np.astype(np.int32)
    nz = np.nonzero(diff)
    perm = np.random.RandomState(3).permutation(len(nz[0]))
    keep = diff.astype(np.uint8)
    drop = perm[: int(len(nz[0]) * 0.9)]
    keep[nz[0][drop], nz[1][drop], nz[2][drop]] = 0
    nxt = torch.empty_like(mask)
    sb.generate_fuzzy_boundaries(ops, halo, nxt, dev(keep).view(-1))
    np.testing.assert_array_equal(nxt.cpu().numpy().reshape(shape), OA.fuzzy_iteration(halo_np, perm))
    # L1 distance thresholds == repeated 6-neighbour dilations of the halo
    d16, t16 = ops.u16("d0"), ops.u16("d1")
    ops.dist(halo, d16, t16, 4, 1)
    d = d16.cpu().numpy().view(np.uint16).reshape(shape)
    want = halo_np
    for kk in range(1, 4):
        want = OA.build_halo(want, 1)
        np.testing.assert_array_equal((d <= kk).astype(np.uint8), want)
