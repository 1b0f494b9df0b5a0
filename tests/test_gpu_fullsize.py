"""Parity at BASELINE.json's full size (256^3) through size-independent properties — the oracle is
too slow there, so these check what must hold for any input: identities, symmetries, linearity,
partition of unity, per-label moments, and agreement of independent kernel variants."""
import numpy as np
import pytest
import torch

from fetalsyngen_b200 import _lib
from fetalsyngen_b200.engine import SamplePlan, engine_for
from fetalsyngen_b200.utils.phantom import label_phantom
from gpu_util import DEV, TOL

pytestmark = pytest.mark.gpu
S = 256
SHAPE = (S, S, S)


@pytest.fixture(scope="module")
def data():
    seg_h, seeds_h = label_phantom(SHAPE)
    eng = engine_for(DEV, SHAPE, (0.5, 0.5, 0.5))
    return {"eng": eng, "seg_h": seg_h, "seeds_h": seeds_h, "seg": torch.from_numpy(seg_h).to(DEV).view(-1),
            "seeds": [torch.from_numpy(s).to(DEV).view(-1) for s in seeds_h]}


def _plan(rs, **kw):
    p = SamplePlan(mus=(25 + 200 * rs.rand(50)).astype(np.float32), sigmas=(5 + 20 * rs.rand(50)).astype(np.float32), rng_seed=int(rs.randint(1 << 30)), sample_id=int(rs.randint(1 << 30)))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _identity(p, with_field=False):
    p.deform, p.A, p.c2, p.center = True, np.eye(3, dtype=np.float32), (np.array(SHAPE) - 1) / 2, ((np.array(SHAPE) - 1) / 2).astype(np.float32)
    if with_field:
        p.fsmall = np.zeros((12, 12, 12, 3), np.float32)
    return p


def test_gmm_per_label_moments_at_full_size(data):
    """Philox mode: mean / std of every label's intensities match (mu, sigma) of the drawn table
    (the north star's validation of the counter-based noise)."""
    eng, rs = data["eng"], np.random.RandomState(0)
    p = _plan(rs)
    out = torch.empty((1, eng.nvox), dtype=torch.float32, device=DEV)
    lab = torch.empty((1, eng.nvox), dtype=torch.uint8, device=DEV)
    eng.gmm([p], [data["seeds"]], out, labels_out=lab)
    want_lab = sum(s.astype(np.int64) for s in data["seeds_h"]).reshape(-1)
    assert np.array_equal(lab[0].cpu().numpy(), want_lab.astype(np.uint8))  # integer work: bit-exact seed sum
    x, l = out[0].double(), lab[0].long()
    for lbl in np.unique(want_lab):
        v = x[l == int(lbl)]
        n = v.numel()
        if n < 20000:
            continue
        mu, sg = float(p.mus[lbl]), float(p.sigmas[lbl])
        if mu - 4 * sg > 0:  # the clamp at 0 does not bite: plain Gaussian moments
            assert abs(v.mean().item() - mu) < 5 * sg / np.sqrt(n) + 1e-3
            assert abs(v.std().item() - sg) < 5 * sg / np.sqrt(2 * n) + 1e-3
        assert v.min().item() >= 0


def test_identity_deformation_returns_the_input(data):
    """A = I, zero control grid: the segmentation comes back bit-exactly, the image exactly except the
    planes the reference's sampler zeroes (coordinate 0 is 'not ok', Appendix A.8)."""
    eng, rs = data["eng"], np.random.RandomState(1)
    img = torch.rand((1, eng.nvox), device=DEV) * 300
    for with_field in (False, True):  # generic kernel / fast path
        p = _identity(_plan(rs), with_field)
        dst, dseg = torch.empty_like(img), torch.empty((1, eng.nvox), dtype=torch.uint8, device=DEV)
        eng.warp([p], img, [data["seg"]], dst, dseg, epilogue=False)
        assert torch.equal(dseg[0], data["seg"])
        a, b = dst[0].view(SHAPE), img[0].view(SHAPE)
        if with_field:
            # fast path: the floor index is clamped to S-2, so on the last planes the corner comes back as
            # a + 1*(b - a), equal to b to an ulp (image path: tolerance), exact everywhere else
            assert torch.equal(a[1:-1, 1:-1, 1:-1], b[1:-1, 1:-1, 1:-1])
            assert float((a[1:, 1:, 1:] - b[1:, 1:, 1:]).abs().max()) <= 300 * 2e-7
        else:
            assert torch.equal(a[1:, 1:, 1:], b[1:, 1:, 1:])
        assert float(a[0].abs().max()) == 0 and float(a[:, 0].abs().max()) == 0 and float(a[:, :, 0].abs().max()) == 0


def test_flip_is_an_involution_on_the_segmentation(data):
    eng, rs = data["eng"], np.random.RandomState(2)
    p = _identity(_plan(rs, flip=True), with_field=True)
    img = torch.rand((1, eng.nvox), device=DEV)
    d1, s1 = torch.empty_like(img), torch.empty((1, eng.nvox), dtype=torch.uint8, device=DEV)
    eng.warp([p], img, [data["seg"]], d1, s1, epilogue=False)
    assert torch.equal(s1[0].view(SHAPE), torch.flip(data["seg"].view(SHAPE), [0]))
    d2, s2 = torch.empty_like(img), torch.empty_like(s1)
    eng.warp([p], d1, [s1[0]], d2, s2, epilogue=False)
    assert torch.equal(s2[0], data["seg"])


def test_fast_and_generic_warp_kernels_agree_at_full_size(data):
    """The generic kernel (forced by a second image channel) and the packed fast path compute the same
    coordinates: identical segmentation, image to tolerance; histogram of labels is plausible."""
    eng, rs = data["eng"], np.random.RandomState(3)
    from fetalsyngen_b200.tables import make_affine_matrix

    p = _plan(rs)
    p.deform, p.flip = True, True
    p.A = make_affine_matrix(np.array([0.3, -0.25, 0.2]), np.array([0.015, -0.01, 0.02]), np.array([1.08, 0.93, 1.02])).astype(np.float32)
    p.c2, p.center = (np.array(SHAPE) - 1) / 2, ((np.array(SHAPE) - 1) / 2).astype(np.float32)
    p.fsmall = (3.0 * rs.randn(13, 11, 15, 3)).astype(np.float32)
    p.gamma, p.bf_low = 1.07, (0.2 * rs.randn(4, 3, 5)).astype(np.float32)
    img = torch.rand((1, eng.nvox), device=DEV) * 300
    outs = []
    for second in (False, True):
        dst, dseg = torch.empty_like(img), torch.empty((1, eng.nvox), dtype=torch.uint8, device=DEV)
        d2 = torch.empty_like(img) if second else None
        eng.warp([p], img, [data["seg"]], dst, dseg, [img[0]] if second else None, d2)
        outs.append((dst.clone(), dseg.clone()))
    assert torch.equal(outs[0][1], outs[1][1])
    rng = float(outs[1][0].max() - outs[1][0].min())
    assert float((outs[0][0] - outs[1][0]).abs().max()) / rng <= TOL
    frac_fg = float((outs[0][1] > 0).float().mean())
    assert 0.5 * float((data["seg"] > 0).float().mean()) < frac_fg < 1.5 * float((data["seg"] > 0).float().mean())


def test_blur_resample_is_linear_and_preserves_constants(data):
    eng, rs = data["eng"], np.random.RandomState(4)
    from fetalsyngen_b200.tables import resample_stds

    p = _plan(rs)
    p.spacing = np.array([1.1] * 3)
    p.stds = resample_stds(p.spacing, [0.5] * 3, 0.4)
    x = torch.rand((1, eng.nvox), device=DEV)
    y = torch.rand((1, eng.nvox), device=DEV)
    t1, t2 = torch.empty_like(x), torch.empty_like(x)

    def run(src):
        dst = torch.empty_like(src)
        info = eng.sepconv([p], src, dst, t1, t2)
        n = int(np.prod(info[0][0]))
        return dst[0, :n].clone(), info[0][0]

    fx, n3 = run(x)
    fy, _ = run(y)
    fxy, _ = run(2.5 * x - 0.75 * y)
    assert float((fxy - (2.5 * fx - 0.75 * fy)).abs().max()) <= 5e-5
    ones, _ = run(torch.ones_like(x))
    v = ones.view(n3)
    m = 12  # away from the zero-padded borders the composed operator is a partition of unity
    assert float((v[m:-m, m:-m, m:-m] - 1).abs().max()) <= 1e-5
    assert float(v.max()) <= 1 + 1e-5 and float(v.min()) >= 0


def test_zoom_normalisations_hit_their_targets(data):
    eng, rs = data["eng"], np.random.RandomState(5)
    n = (117, 117, 117)
    src = (torch.rand(int(np.prod(n)), device=DEV) * 200 + 3).contiguous()
    fac = [SHAPE[a] / n[a] for a in range(3)]
    outs = {}
    for post in (0, 1, 2):
        dst = torch.empty((1, eng.nvox), dtype=torch.float32, device=DEV)
        eng.zoom([src], [n], [fac], dst, post=post)
        outs[post] = dst[0].clone()
    raw = outs[0]
    assert float(raw.max()) <= float(src.max()) + 1e-3 and float(raw.min()) >= float(src.min()) - 1e-3  # convex combinations
    assert float(outs[1].max()) == 1.0 and float((outs[1] - raw / raw.max()).abs().max()) <= 2e-6
    # the fused (v * a + b) form maps the maximum to exactly 1; the minimum to 0 within an ulp of the range
    # (exactly 0 when the minimum is 0, which generated images always have outside the field of view)
    assert float(outs[2].max()) == 1.0 and 0.0 <= float(outs[2].min()) <= 1e-7
    want = (raw - raw.min()) / (raw.max() - raw.min())
    assert float((outs[2] - want).abs().max()) <= 5e-6
    const = torch.full_like(src, 7.0)
    dst = torch.empty((1, eng.nvox), dtype=torch.float32, device=DEV)
    eng.zoom([const], [n], [fac], dst, post=0)
    assert float((dst - 7.0).abs().max()) <= 1e-5


def test_batch_entries_are_independent_at_full_size(data):
    """A batch of 4 equals four single launches bit for bit (job structs do not leak into each other)."""
    eng, rs = data["eng"], np.random.RandomState(6)
    from fetalsyngen_b200.tables import make_affine_matrix, resample_stds

    plans = []
    for i in range(4):
        p = _plan(rs)
        p.deform, p.flip = True, bool(i % 2)
        p.A = make_affine_matrix((rs.rand(3) - 0.5) * 0.6, (rs.rand(3) - 0.5) * 0.04, 1 + (rs.rand(3) - 0.5) * 0.2).astype(np.float32)
        p.c2, p.center = (np.array(SHAPE) - 1) / 2, ((np.array(SHAPE) - 1) / 2).astype(np.float32)
        s = int(rs.randint(8, 16))
        p.fsmall = (2.0 * rs.randn(s, s, s, 3)).astype(np.float32)
        p.gamma, p.bf_low = float(np.exp(0.1 * rs.randn())), (0.2 * rs.randn(3, 3, 3)).astype(np.float32)
        p.spacing = np.array([0.5 + rs.rand()] * 3)
        p.stds = resample_stds(p.spacing, [0.5] * 3, rs.rand())
        p.noise_std = 8.0
        plans.append(p)
    img_b, seg_b = eng.run_base(plans, [data["seeds"]] * 4, [data["seg"]] * 4, scale=True)
    img_b, seg_b = img_b.clone(), seg_b.clone()
    for i, p in enumerate(plans):
        im, sg = eng.run_base([p], [data["seeds"]], [data["seg"]], scale=True)
        assert torch.equal(im[0], img_b[i]) and torch.equal(sg[0], seg_b[i])
        assert float(im.min()) == 0.0 and float(im.max()) == 1.0


# ----------------------------------------------------------------------------- seed generation and packed cache at 256^3
def test_seed_generation_at_full_size(data):
    """scripts/generate_seeds.py for one 256^3 subject through properties that hold for any input: the
    partition covers every labelled voxel once, every fit converges with ordered lower bounds, the seed
    volumes keep the meta-label support with labels 10 m .. 10 m + n - 1 and equal the oracle's argmax
    under the fitted parameters, and the whole run is reproducible for a fixed seed."""
    from fetalsyngen_b200.seeds import SeedGenerator

    seg = data["seg_h"]
    rs = np.random.RandomState(3)
    base = np.array([0, 900, 300, 450, 820, 520, 330, 480], dtype=np.float32)[seg]
    image = np.maximum(base + 45 * rs.standard_normal(SHAPE).astype(np.float32), 0)
    image[(seg == 0) & (rs.rand(*SHAPE) < 0.7)] = 0  # most of the background is air
    gen = SeedGenerator("feta", DEV, seed=12)
    x, index, counts = gen.partition(image, seg)
    meta = np.array([4, 1, 2, 3, 1, 3, 2, 3], dtype=np.uint8)[seg]
    meta[(seg == 0) & (image == 0)] = 0
    assert counts == [int((meta == m).sum()) for m in range(1, 5)]
    out = gen.split_labels(image, seg, [1, 3, 5])
    fits = gen.last_fit
    for (n_sub, m), f in fits.items():
        assert f["converged"] and 2 <= f["n_iter"] <= 100
        assert np.all(np.diff(f["trace"][1:]) > -1e-6)  # EM never lowers the likelihood (first step starts from point masses)
        assert abs(f["weights"].sum() - 1) < 1e-12 and np.all(f["covariances"] > 0)
    for n_sub in (1, 3, 5):
        for m in range(1, 5):
            vol = out[n_sub][m].cpu().numpy()
            sel = meta == m
            assert not vol[~sel].any()
            labs = np.unique(vol[sel])
            assert labs.min() >= 10 * m and labs.max() <= 10 * m + n_sub - 1
            if n_sub > 1:
                # the labels are the argmax of the fitted mixture's weighted log-probabilities (oracle arithmetic, subsample)
                import np_seeds as NS

                pick = np.flatnonzero(sel.reshape(-1))[:: max(1, int(sel.sum()) // 50000)]
                want = NS.predict(image.reshape(-1)[pick], fits[(n_sub, m - 1)]) + 10 * m
                assert (vol.reshape(-1)[pick] != want).mean() <= 1e-4
    again = SeedGenerator("feta", DEV, seed=12).split_labels(image, seg, [3])
    assert all(torch.equal(again[3][m], out[3][m]) for m in range(1, 5))


def test_packed_seed_cache_at_full_size(data):
    """Six sub-class counts of a 256^3 subject in one uint16 volume: every draw of counts unpacks to the
    sum of the four selected seed volumes, and the GMM image from the unpacked labels equals the one from
    the four volumes (same Philox stream)."""
    from fetalsyngen_b200.data import packed as K

    seeds = {}
    for n in range(1, 7):
        _, sv = label_phantom(SHAPE, n_sub=(n, n, n, n), seed=n)
        seeds[n] = {m + 1: sv[m] for m in range(4)}
    words, counts = K.pack_seed_volumes(seeds)
    assert words.dtype == np.uint16 and words.nbytes == 2 * S**3
    ps = K.PackedSeeds(words, counts, DEV)
    eng = data["eng"]
    rs = np.random.RandomState(8)
    for _ in range(3):
        m2s = {m: int(rs.randint(1, 7)) for m in range(1, 5)}
        lab = ps.labels(m2s, DEV)
        vols = [torch.from_numpy(seeds[m2s[m]][m]).to(DEV).view(-1) for m in range(1, 5)]
        want = sum(v.to(torch.int32) for v in vols).to(torch.uint8)
        assert torch.equal(lab.view(-1), want)
        p = _plan(rs)
        a = torch.empty((1, S**3), dtype=torch.float32, device=DEV)
        b = torch.empty_like(a)
        eng.gmm([p], [vols], a)
        eng.gmm([p], [[lab.view(-1)]], b)
        assert torch.equal(a, b)
