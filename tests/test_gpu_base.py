"""GPU parity tests of the base generation path: CUDA kernels (through the C-ABI) vs the
golden vectors of the unmodified reference and vs the numpy oracle."""
import numpy as np
import pytest
import torch

import np_oracle as O
from golden_util import BASE_CASES, load_case, oracle_params

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from fetalsyngen_b200 import _lib
    from fetalsyngen_b200.engine import SamplePlan, engine_for
    from gpu_util import DEV, TOL, engine_from_golden, plan_from_golden, rel_err, seeds_from_golden


def test_library_loaded_and_abi():
    lib = _lib.load(build_if_missing=False)
    assert lib.fsg_version() == 102


def test_philox_raw_matches_published_algorithm():
    n = 4096
    out = torch.empty(n, dtype=torch.float32, device=DEV)
    rng = _lib.Rng(0x299F31D0A4093822, 0x0370734413198A2E, 0x85A308D3, 0)
    _lib.call("fsg_philox_fill", rng, out.data_ptr(), n, 1, torch.cuda.current_stream().cuda_stream)
    words = out.cpu().numpy().view(np.uint32)
    for blk in (0, 1, 7, 1023):
        want = O.philox4x32_10([blk, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])
        assert list(words[4 * blk : 4 * blk + 4]) == want
    # Random123 KAT (counter word 0 = 0x243f6a88 is block index 0x243f6a88: check via the oracle at block 0 instead)
    assert O.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_normals_moments_and_independence():
    n = 1 << 22
    out = torch.empty(n, dtype=torch.float32, device=DEV)
    _lib.call("fsg_philox_fill", _lib.Rng(1234, 5, 1, 0), out.data_ptr(), n, 0, torch.cuda.current_stream().cuda_stream)
    x = out.double()
    assert abs(x.mean().item()) < 4 / np.sqrt(n)
    assert abs(x.std().item() - 1) < 3e-3
    assert abs((x**3).mean().item()) < 0.02 and abs((x**4).mean().item() - 3) < 0.05
    # lag-1..4 autocorrelation ~ 0 (covers the 4 outputs of one Philox block)
    for lag in (1, 2, 3, 4, 256):
        assert abs((x[:-lag] * x[lag:]).mean().item()) < 5 / np.sqrt(n)
    # Kolmogorov-Smirnov against the normal CDF
    xs = np.sort(out.cpu().numpy()[: 1 << 18].astype(np.float64))
    from math import erf

    cdf = 0.5 * (1 + np.vectorize(erf)(xs / np.sqrt(2)))
    ks = np.abs(cdf - (np.arange(1, xs.size + 1) / xs.size)).max()
    assert ks < 1.95 / np.sqrt(xs.size)
    # a different stream is a different sequence
    out2 = torch.empty(1024, dtype=torch.float32, device=DEV)
    _lib.call("fsg_philox_fill", _lib.Rng(1234, 6, 1, 0), out2.data_ptr(), 1024, 0, torch.cuda.current_stream().cuda_stream)
    assert not torch.equal(out[:1024], out2)


@pytest.mark.parametrize("name", BASE_CASES)
def test_gmm_bit_exact(name):
    d = load_case(name)
    eng = engine_from_golden(d)
    plan = plan_from_golden(d)
    out = torch.empty((1, eng.nvox), dtype=torch.float32, device=DEV)
    lab = torch.empty((1, eng.nvox), dtype=torch.uint8, device=DEV)
    eng.gmm([plan], [seeds_from_golden(d)], out, labels_out=lab)
    np.testing.assert_array_equal(lab.cpu().numpy().reshape(d["labels"].shape), d["labels"])
    want = d["intensity"] if "intensity" in d else O.gmm_intensities(d["labels"], d["mus"], d["sigmas"], d["gmm_noise"])
    np.testing.assert_array_equal(out.cpu().numpy().reshape(want.shape), want)


@pytest.mark.parametrize("name", [c for c in BASE_CASES if c != "c32_gates_off"])
def test_warp_coordinates_bit_exact(name):
    d = load_case(name)
    eng = engine_from_golden(d)
    plan = plan_from_golden(d)
    got = eng.warp_coords(plan).cpu().numpy()
    if "coords" in d:
        want = d["coords"]
    else:
        p = oracle_params(d)
        fs = p.get("Fsmall")
        F = None if fs is None else O.zoom_linear(fs, np.asarray(d["labels"].shape) / np.asarray(fs.shape[:3]))
        want = np.stack(O.deformation_coords(d["labels"].shape, d["labels"].shape, d["A"], d["c2"], F))
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("name", BASE_CASES)
@pytest.mark.parametrize("scale", [False, True])
def test_base_pipeline_vs_reference_golden(name, scale):
    d = load_case(name)
    eng = engine_from_golden(d)
    plan = plan_from_golden(d)
    seg = torch.from_numpy(d["seg_in"]).to(DEV).contiguous().view(-1)
    img, sg = eng.run_base([plan], [seeds_from_golden(d)], [seg], scale=scale)
    np.testing.assert_array_equal(sg[0].cpu().numpy(), d["seg_out"])  # bit exact
    want = d["scaled"] if scale else d["final"]
    assert rel_err(img[0], want) <= TOL


@pytest.mark.parametrize("name", [c for c in BASE_CASES if "intensity" in load_case(c)])
def test_stage_outputs_vs_reference_golden(name):
    d = load_case(name)
    eng = engine_from_golden(d)
    plan = plan_from_golden(d)
    shape = d["labels"].shape
    seg = torch.from_numpy(d["seg_in"]).to(DEV).contiguous().view(1, -1)
    inten = torch.from_numpy(d["intensity"]).to(DEV).contiguous().view(1, -1)
    # warp + gamma + bias (fused epilogue) vs the reference's bias_out
    dst = torch.empty_like(inten)
    dseg = torch.empty_like(seg)
    eng.warp([plan], inten, seg, dst, dseg)
    assert rel_err(dst.view(shape), d["bias_out"]) <= TOL
    # gamma alone and bias alone (identity-mode launches)
    _, _, st = O.generate_base(d["labels"], d["seg_in"], oracle_params(d))
    warped = torch.from_numpy(st["warped"]).to(DEV).contiguous().view(1, -1)
    g = torch.empty_like(warped)
    eng.warp([SamplePlan(gamma=plan.gamma)], warped, None, g, None)
    assert rel_err(g.view(shape), d["gamma_out"]) <= TOL
    # blur
    if "blurred" in d:
        src = torch.from_numpy(d["bias_out"]).to(DEV).contiguous().view(1, -1)
        bl, tmp = torch.empty_like(src), torch.empty_like(src)
        eng.blur([plan.stds], src, bl, tmp)
        assert rel_err(bl.view(shape), d["blurred"]) <= TOL
        low = torch.empty((1, d["lowres"].size), dtype=torch.float32, device=DEV)
        blurred = torch.from_numpy(d["blurred"]).to(DEV).contiguous().view(1, -1)
        info = eng.resample([plan], blurred, low)
        assert info[0][0] == d["lowres"].shape
        np.testing.assert_allclose(info[0][1], d["factors"], rtol=0, atol=0)
        assert rel_err(low.view(d["lowres"].shape), d["noisy"]) <= TOL
        up = torch.empty((1, eng.nvox), dtype=torch.float32, device=DEV)
        noisy = torch.from_numpy(d["noisy"]).to(DEV).contiguous().view(-1)
        eng.zoom([noisy], [d["noisy"].shape], [1 / d["factors"]], up, post=1)
        assert rel_err(up.view(shape), d["final"]) <= 1e-6
        assert up.max().item() == 1.0


@pytest.mark.parametrize("name", [c for c in BASE_CASES if "blurred" in load_case(c)])
def test_fused_blur_resample_vs_reference_golden(name):
    """fsg_sep_compose + fsg_sepconv (blur composed with the down-sampling, noise epilogue) against
    the reference's blur -> interp -> noise outputs."""
    d = load_case(name)
    eng = engine_from_golden(d)
    plan = plan_from_golden(d)
    shape = d["labels"].shape
    src = torch.from_numpy(d["bias_out"]).to(DEV).contiguous().view(1, -1)
    low, tmp = torch.empty_like(src), torch.empty_like(src)
    info = eng.sepconv([plan], src, low, low, tmp)
    assert info[0][0] == d["lowres"].shape
    np.testing.assert_allclose(info[0][1], d["factors"], rtol=0, atol=0)
    got = low[0, : d["noisy"].size].view(d["noisy"].shape)
    assert rel_err(got, d["noisy"]) <= TOL
    # the same kernels with identity positions are the plain separable blur
    bl = torch.empty_like(src)
    tmp1 = torch.empty_like(src)
    eng.sepconv([SamplePlan(stds=plan.stds)], src, bl, tmp1, tmp, positions=False)
    assert rel_err(bl.view(shape), d["blurred"]) <= TOL


@pytest.mark.parametrize("shape,sigma", [((40, 36, 44), 0.6), ((40, 36, 44), 2.4), ((31, 45, 38), 4.9), ((64, 48, 80), 11.0)])
def test_sepconv_blur_all_window_widths(shape, sigma):
    """Window widths 5..67 taps: the W=8/16/32 register kernels and the generic fallback."""
    rs = np.random.RandomState(int(sigma * 10))
    x = (rs.rand(*shape) * 300).astype(np.float32)
    stds = np.array([sigma, sigma * 0.5, sigma])
    want = O.gaussian_blur_3d(x, stds)
    eng = engine_for(DEV, shape, (0.5, 0.5, 0.5))
    src = torch.from_numpy(x).to(DEV).view(1, -1)
    out, t1, t2 = torch.empty_like(src), torch.empty_like(src), torch.empty_like(src)
    eng.sepconv([SamplePlan(stds=stds)], src, out, t1, t2, positions=False)
    assert rel_err(out.view(shape), want) <= 1e-5


def test_sepconv_philox_noise_moments():
    """Philox-mode noise of the fused kernel: N(0, std) residuals, clamp at 0, fresh per sample."""
    shape = (96, 96, 96)
    eng = engine_for(DEV, shape, (0.5, 0.5, 0.5))
    src = torch.full((2, eng.nvox), 200.0, dtype=torch.float32, device=DEV)
    plans = []
    for sid in (1, 2):
        p = SamplePlan(spacing=np.array([0.8] * 3), stds=np.array([0.0] * 3), noise_std=10.0, rng_seed=77, sample_id=sid)
        plans.append(p)
    low, tmp = torch.empty_like(src), torch.empty_like(src)
    info = eng.sepconv(plans, src, low, low, tmp)
    n = int(np.prod(info[0][0]))
    a, b = low[0, :n].double(), low[1, :n].double()
    # interior voxels are exactly 200 before the noise (weights sum to 1); borders are partly zero padded
    vol = a.view(info[0][0])[2:-2, 2:-2, 2:-2].reshape(-1)
    assert abs(vol.mean().item() - 200) < 0.05 and abs(vol.std().item() - 10) < 0.05
    r = (vol - 200) / 10
    assert abs((r**3).mean().item()) < 0.03 and abs((r**4).mean().item() - 3) < 0.08
    for lag in (1, 2, 3, 4, info[0][0][2]):
        assert abs((r[:-lag] * r[lag:]).mean().item()) < 5 / np.sqrt(r.numel())
    assert not torch.equal(a, b) and (low[:, :n] >= 0).all()


def _random_plan(rs, shape, dev, deform=True, resample=True):
    p = SamplePlan()
    p.mus = (25 + 200 * rs.rand(50)).astype(np.float32)
    p.sigmas = (5 + 20 * rs.rand(50)).astype(np.float32)
    n = int(np.prod(shape))
    p.gmm_noise = torch.from_numpy(rs.randn(n).astype(np.float32)).to(dev)
    p.flip = bool(rs.rand() < 0.5)
    if deform:
        from fetalsyngen_b200.tables import make_affine_matrix

        rot = (2 * 20 * rs.rand(3) - 20) / 180 * np.pi
        p.deform = True
        p.A = make_affine_matrix(rot, 0.04 * rs.rand(3) - 0.02, 1 + 0.2 * rs.rand(3) - 0.1).astype(np.float32)
        p.c2 = (np.array(shape) - 1) / 2
        p.center = ((np.array(shape) - 1) / 2).astype(np.float32)
        s = [max(int(round(0.045 * v * (1 + 0.3 * rs.rand()))), 1) for v in shape]
        p.fsmall = (3.0 * rs.randn(*s, 3)).astype(np.float32)
    p.gamma = float(np.exp(0.1 * rs.randn()))
    b = [max(int(round(0.012 * v)), 1) for v in shape]
    p.bf_low = (0.2 * rs.randn(*b)).astype(np.float32)
    if resample:
        sp = 0.5 + rs.rand()
        p.spacing = np.array([sp] * 3)
        from fetalsyngen_b200.tables import resample_stds

        p.stds = resample_stds(p.spacing, [0.5] * 3, rs.rand())
        p.noise_std = float(5 + 10 * rs.rand())
        from fetalsyngen_b200.tables import resample_size

        m = int(np.prod([resample_size(v, 0.5, sp) for v in shape]))
        p.noise = torch.from_numpy(rs.randn(m).astype(np.float32)).to(dev)
    return p


def _plan_to_oracle(p):
    q = {"mus": p.mus, "sigmas": p.sigmas, "gmm_noise": p.gmm_noise.cpu().numpy(), "flip": p.flip, "resolution": np.array([0.5] * 3)}
    if p.deform:
        q.update(A=p.A, c2=p.c2, Fsmall=p.fsmall)
    q.update(gamma=p.gamma, bf_low=p.bf_low)
    if p.spacing is not None:
        q.update(spacing=p.spacing, stds=p.stds, noise_std=p.noise_std, noise=None)
    return q


def _phantom(rs, shape):
    """Nested-ellipsoid label phantom: seg labels 0..7 and seed labels 0,10..49."""
    g = np.meshgrid(*[np.linspace(-1, 1, s) for s in shape], indexing="ij")
    r = np.sqrt(sum((gi / (0.55 + 0.1 * a)) ** 2 for a, gi in enumerate(g)))
    seg = np.digitize(r, [0.25, 0.45, 0.6, 0.75, 0.85, 0.93, 1.0][::1])
    seg = (7 - seg).clip(0, 7).astype(np.uint8)
    sub = rs.randint(0, 4, size=shape)
    meta = np.array([0, 1, 1, 2, 3, 3, 4, 4])[seg]
    lab = np.where(meta > 0, meta * 10 + sub, 0).astype(np.int8)
    seeds = [np.where(meta == m, lab, 0).astype(np.int8) for m in range(1, 5)]
    return seg, seeds


@pytest.mark.parametrize("shape,seed", [((64, 64, 64), 0), ((72, 56, 80), 1), ((96, 96, 96), 2), ((33, 47, 29), 3)])
def test_base_pipeline_vs_oracle_random(shape, seed):
    rs = np.random.RandomState(seed)
    seg, seeds = _phantom(rs, shape)
    p = _random_plan(rs, shape, DEV)
    eng = engine_for(DEV, shape, (0.5, 0.5, 0.5))
    q = _plan_to_oracle(p)
    q["noise"] = p.noise.cpu().numpy().reshape(eng.lowres_shape(p.spacing))
    q["gmm_noise"] = q["gmm_noise"].reshape(shape)
    want_img, want_seg, _ = O.generate_base(sum(s.astype(np.int64) for s in seeds), seg, q)
    dseeds = [torch.from_numpy(s).to(DEV).view(-1) for s in seeds]
    img, sg = eng.run_base([p], [dseeds], [torch.from_numpy(seg).to(DEV).view(-1)])
    np.testing.assert_array_equal(sg[0].cpu().numpy(), want_seg)
    assert rel_err(img[0], want_img) <= TOL


@pytest.mark.parametrize("shape,res,spacing", [((48, 40, 56), 1.0, 0.6), ((33, 47, 29), 1.0, 0.75), ((64, 64, 64), 0.8, 0.5)])
def test_resolution_simulation_with_spacing_finer_than_the_resolution(shape, res, spacing):
    """spacing < resolution: the reference zeroes the blur and UP-samples (synthseg.py:78-84), so the coarse grid
    outgrows the volume (n_out > n_in on every axis).  Both schedules (z+y fused: extents divisible by 4; generic:
    odd extents) against the oracle, in a batch next to an ordinary down-sampling sample whose buffers follow
    directly behind (an overrun of the scratch rows would corrupt it)."""
    from fetalsyngen_b200.tables import resample_size, resample_stds

    rs = np.random.RandomState(int(spacing * 100))
    seg, seeds = _phantom(rs, shape)
    eng = engine_for(DEV, shape, (res, res, res))
    plans, want = [], []
    for sp in (spacing, 1.7 * res):
        p = _random_plan(rs, shape, DEV)
        p.spacing = np.array([sp] * 3)
        p.stds = resample_stds(p.spacing, [res] * 3, rs.rand())
        n = [resample_size(v, res, sp) for v in shape]
        assert (sp >= res) or all(n[a] > shape[a] for a in range(3))
        p.noise = torch.from_numpy(rs.randn(int(np.prod(n))).astype(np.float32)).to(DEV)
        q = _plan_to_oracle(p)
        q["resolution"] = np.array([res] * 3)
        q["noise"] = p.noise.cpu().numpy().reshape(n)
        q["gmm_noise"] = q["gmm_noise"].reshape(shape)
        plans.append(p)
        want.append(O.generate_base(sum(s_.astype(np.int64) for s_ in seeds), seg, q)[:2])
    dseeds = [torch.from_numpy(s_).to(DEV).view(-1) for s_ in seeds]
    dseg = torch.from_numpy(seg).to(DEV).view(-1)
    img, sg = eng.run_base(plans, [dseeds] * 2, [dseg] * 2)
    for b in range(2):
        np.testing.assert_array_equal(sg[b].cpu().numpy(), want[b][1])
        assert rel_err(img[b], want[b][0]) <= TOL
    # the per-sample API (FetalSynthGen.augment's tail) sizes its own scratch
    too_small = torch.empty((1, eng.nvox), dtype=torch.float32, device=DEV)
    with pytest.raises((ValueError, Exception)):
        eng.sepconv([plans[0]], too_small.clone(), too_small, too_small, too_small.clone())


def test_batched_launch_equals_single_launches():
    shape = (48, 40, 56)
    rs = np.random.RandomState(5)
    seg, seeds = _phantom(rs, shape)
    eng = engine_for(DEV, shape, (0.5, 0.5, 0.5))
    plans = [_random_plan(rs, shape, DEV, deform=(i != 2), resample=(i != 1)) for i in range(4)]
    plans[3].noise_std, plans[3].noise = None, None
    dseeds = [torch.from_numpy(s).to(DEV).view(-1) for s in seeds]
    dseg = torch.from_numpy(seg).to(DEV).view(-1)
    img_b, seg_b = eng.run_base(plans, [dseeds] * 4, [dseg] * 4)
    img_b, seg_b = img_b.clone(), seg_b.clone()
    for i, p in enumerate(plans):
        im, sg = eng.run_base([p], [dseeds], [dseg])
        assert torch.equal(im[0], img_b[i]) and torch.equal(sg[0], seg_b[i])


def test_warp_kernel_variants_agree(monkeypatch):
    """The three warp kernels (generic, full-z fast path, TMA-staged tiles) must give the same
    segmentation bit for bit and the same image to tolerance on a 64^3 case."""
    import os

    d = load_case("c64_default")
    eng = engine_from_golden(d)
    plan = plan_from_golden(d)
    seg = torch.from_numpy(d["seg_in"]).to(DEV).contiguous().view(-1)
    img0, sg0 = eng.run_base([plan], [seeds_from_golden(d)], [seg], scale=False)
    np.testing.assert_array_equal(sg0[0].cpu().numpy(), d["seg_out"])
    # the TMA variant is chosen by an environment variable read once per process: run it in a child
    import subprocess
    import sys

    code = (
        "import sys; sys.path[:0]=['.','oracle','tests']\n"
        "import numpy as np, torch\n"
        "from golden_util import load_case\n"
        "from gpu_util import plan_from_golden, seeds_from_golden, engine_from_golden, rel_err\n"
        "d=load_case('c64_default'); eng=engine_from_golden(d); plan=plan_from_golden(d)\n"
        "seg=torch.from_numpy(d['seg_in']).cuda().contiguous().view(-1)\n"
        "img,sg=eng.run_base([plan],[seeds_from_golden(d)],[seg],scale=False)\n"
        "assert np.array_equal(sg[0].cpu().numpy(), d['seg_out'])\n"
        "print('ERR', rel_err(img[0], d['final']))\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, env=dict(os.environ, FSG_WARP_TILE="1", FSG_TILE_DEBUG="2"))
    assert res.returncode == 0, res.stderr[-1500:]
    assert "tile (" not in res.stdout  # no box-bound violations reported by the debug check
    err = float(res.stdout.split("ERR")[1])
    assert err <= TOL
    # the pipelined TMA kernel (persistent blocks, producer / consumer warps): staged boxes, then every tile forced
    # through its global-memory fallback (FSG_TILE_DEBUG=1)
    for extra in ({}, {"FSG_TILE_DEBUG": "1"}):
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, env=dict(os.environ, FSG_WARP_PIPE="1", FSG_WARP_TILE="0", **extra), timeout=300)
        assert res.returncode == 0, res.stderr[-1500:]
        assert float(res.stdout.split("ERR")[1]) <= TOL
    # the fused x + y kernel of the resolution simulation (opt-in: slower than the two streaming passes)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, env=dict(os.environ, FSG_SEP_XY="1"), timeout=300)
    assert res.returncode == 0, res.stderr[-1500:]
    assert float(res.stdout.split("ERR")[1]) <= TOL
    # the linear float32 hand-over (FSG_WARP_TEX=0; the default gathers from a block-linear texture volume)
    if os.environ.get("FSG_WARP_TEX", "1") != "0":
        assert eng.use_tex and eng.tex_eligible(plan)
        eng.use_tex = False
        try:
            img1, sg1 = eng.run_base([plan], [seeds_from_golden(d)], [seg], scale=False)
        finally:
            eng.use_tex = True
        assert torch.equal(img0, img1) and torch.equal(sg0, sg1)


def test_texture_volume_round_trip_and_non_power_of_two_extents():
    """fsg_texvol: linear -> array -> linear is the identity; fsg_gmm's surface output equals its linear output
    on extents whose row and plane counts are not powers of two (the kernel's division path)."""
    from fetalsyngen_b200.engine import TexVolume

    rs = np.random.RandomState(4)
    for shape in [(40, 36, 44), (8, 4, 4), (16, 100, 12)]:
        n = int(np.prod(shape))
        x = torch.from_numpy(rs.randn(n).astype(np.float32)).to(DEV)
        tv = TexVolume(shape)
        tv.upload(x)
        y = torch.empty_like(x)
        tv.download(y)
        torch.cuda.synchronize()
        assert torch.equal(x, y)
        seg, seeds = _phantom(rs, shape)
        p = _random_plan(rs, shape, DEV)
        eng = engine_for(DEV, shape, (0.5, 0.5, 0.5))
        sd = [[torch.from_numpy(v).to(DEV).view(-1) for v in seeds]]
        lin = torch.empty((1, n), dtype=torch.float32, device=DEV)
        for philox in (False, True):
            if philox:
                p.gmm_noise, p.rng_seed, p.sample_id = None, 3, 1
            eng.gmm([p], sd, lin)
            eng.gmm([p], sd, [None], tex=[tv])
            tv.download(y)
            torch.cuda.synchronize()
            assert torch.equal(lin[0], y), (shape, philox)


def test_sample_batch_streams_do_not_depend_on_the_sharding():
    """sample ids generated as one batch or split over two 'ranks' give bit-identical volumes:
    every draw is a function of (base_seed, sample id) (sharding.py), including the control grids
    drawn on the device."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import bench
    from fetalsyngen_b200.sharding import shard_ids
    from fetalsyngen_b200.utils.phantom import label_phantom

    shape = (64, 64, 64)
    seg_h, seeds_h = label_phantom(shape)
    gen = bench.build_generator(shape, DEV)
    seg_d = torch.from_numpy(seg_h).to(DEV)
    seeds_d = [torch.from_numpy(s).to(DEV) for s in seeds_h]

    def run(ids):
        img, seg, params = gen.sample_batch([seg_d] * len(ids), [seeds_d] * len(ids), scale=True, sample_ids=ids, base_seed=77)
        return {i: (img[n].clone(), seg[n].clone(), params[n]) for n, i in enumerate(ids)}

    whole = run([0, 1, 2, 3])
    parts = {}
    for r in range(2):
        parts.update(run(shard_ids(0, 4, r, 2)))
    for i in range(4):
        assert torch.equal(whole[i][0], parts[i][0]) and torch.equal(whole[i][1], parts[i][1])
        assert whole[i][2]["gamma_params"] == parts[i][2]["gamma_params"]
    assert not torch.equal(whole[0][0], whole[1][0])
    again = run([3])
    assert torch.equal(again[3][0], whole[3][0])
    for i in range(4):  # sanity of the generated data
        v = whole[i][0]
        assert torch.isfinite(v).all() and float(v.min()) == 0.0 and float(v.max()) == 1.0
        assert set(torch.unique(whole[i][1]).tolist()) <= set(range(8))


def test_fixed_point_pairs_handover_stays_within_tolerance():
    """FSG_GMM_PAIRS=1 (opt-in): the GMM -> warp hand-over as 16-bit fixed-point z-pairs keeps the
    segmentation bit-exact and the image within the float tolerance of the reference golden."""
    import os
    import subprocess
    import sys

    code = (
        "import sys; sys.path[:0]=['.','oracle','tests']\n"
        "import numpy as np, torch\n"
        "from golden_util import load_case\n"
        "from gpu_util import plan_from_golden, seeds_from_golden, engine_from_golden, rel_err\n"
        "for name in ('c64_default','c32_all'):\n"
        "    d=load_case(name); eng=engine_from_golden(d); plan=plan_from_golden(d)\n"
        "    assert eng.pairs_eligible(plan)\n"
        "    seg=torch.from_numpy(d['seg_in']).cuda().contiguous().view(-1)\n"
        "    img,sg=eng.run_base([plan],[seeds_from_golden(d)],[seg],scale=False)\n"
        "    assert np.array_equal(sg[0].cpu().numpy(), d['seg_out'])\n"
        "    print('ERR', rel_err(img[0], d['final']))\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, env=dict(os.environ, FSG_GMM_PAIRS="1"))
    assert res.returncode == 0, res.stderr[-1500:]
    errs = [float(l.split()[1]) for l in res.stdout.splitlines() if l.startswith("ERR")]
    assert len(errs) == 2 and max(errs) <= TOL, errs


def test_float_pairs_handover_is_bit_identical():
    """FSG_GMM_PAIRS=2: the GMM -> warp hand-over as float2 z-pairs (I[z], I[z+1]) is lossless — image and
    segmentation equal the plain path bit for bit (Philox noise and injected noise, with and without the
    gamma / bias epilogue), and the reference goldens still hold."""
    import os
    import subprocess
    import sys

    code = (
        "import sys; sys.path[:0]=['.','oracle','tests']\n"
        "import numpy as np, torch\n"
        "from golden_util import load_case\n"
        "from gpu_util import plan_from_golden, seeds_from_golden, engine_from_golden, rel_err\n"
        "import fetalsyngen_b200.engine as E\n"
        "for name in ('c64_default','c32_all','c32_affine_only'):\n"
        "    d=load_case(name); eng=engine_from_golden(d)\n"
        "    seg=torch.from_numpy(d['seg_in']).cuda().contiguous().view(-1)\n"
        "    outs=[]\n"
        "    for mode in (2, 0):\n"
        "        E._PAIRS_MODE, E._PAIRS = mode, mode in (1, 2)\n"
        "        for philox in (False, True):\n"
        "            plan=plan_from_golden(d)\n"
        "            if philox: plan.gmm_noise=None; plan.noise=None; plan.rng_seed=5; plan.sample_id=9\n"
        "            img,sg=eng.run_base([plan],[seeds_from_golden(d)],[seg],scale=False)\n"
        "            outs.append((img[0].clone(), sg[0].clone()))\n"
        "    assert torch.equal(outs[0][0], outs[2][0]) and torch.equal(outs[0][1], outs[2][1]), name\n"
        "    assert torch.equal(outs[1][0], outs[3][0]) and torch.equal(outs[1][1], outs[3][1]), name\n"
        "    assert np.array_equal(outs[0][1].cpu().numpy(), d['seg_out'])\n"
        "    print('ERR', rel_err(outs[0][0], d['final']), eng.pairs_eligible(plan_from_golden(d)))\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, env=dict(os.environ, FSG_GMM_PAIRS="2"))
    assert res.returncode == 0, res.stderr[-1500:]
    errs = [float(l.split()[1]) for l in res.stdout.splitlines() if l.startswith("ERR")]
    assert len(errs) == 3 and max(errs) <= TOL, errs


@pytest.mark.parametrize("shape,seed", [((64, 64, 64), 7), ((40, 52, 36), 8)])
def test_second_image_channel_vs_oracle(shape, seed):
    """load_image=True (datasets.py:279-306, affine_nonrigid.py:190-191): the real image rides through the
    same deformation as the synthetic one (third fast_3D_interp_torch call) — generic warp kernel."""
    rs = np.random.RandomState(seed)
    seg, seeds = _phantom(rs, shape)
    p = _random_plan(rs, shape, DEV, resample=False)
    p.gamma, p.bf_low = None, None
    eng = engine_for(DEV, shape, (0.5, 0.5, 0.5))
    image = (rs.rand(*shape) * 1000).astype(np.float32)
    q = _plan_to_oracle(p)
    q["gmm_noise"] = q["gmm_noise"].reshape(shape)
    q["gamma"], q["bf_low"] = None, None
    lab = sum(s.astype(np.int64) for s in seeds)
    want_out = O.gmm_intensities(lab, q["mus"], q["sigmas"], q["gmm_noise"])
    F = O.zoom_linear(q["Fsmall"], np.asarray(shape, dtype=np.float64) / np.asarray(q["Fsmall"].shape[:3], dtype=np.float64))
    coords = O.deformation_coords(shape, shape, q["A"], q["c2"], F)
    want_out, want_seg, want_img = O.apply_deformation(want_out, seg, coords, q["flip"], image)
    src = torch.empty((1, eng.nvox), dtype=torch.float32, device=DEV)
    dseeds = [torch.from_numpy(s).to(DEV).view(-1) for s in seeds]
    eng.gmm([p], [dseeds], src)
    dst, dseg, dst2 = torch.empty_like(src), torch.empty((1, eng.nvox), dtype=torch.uint8, device=DEV), torch.empty_like(src)
    eng.warp([p], src, [torch.from_numpy(seg).to(DEV).view(-1)], dst, dseg, [torch.from_numpy(image).to(DEV).view(-1)], dst2, epilogue=False)
    np.testing.assert_array_equal(dseg[0].cpu().numpy().reshape(shape), want_seg)
    assert rel_err(dst[0].view(shape), want_out) <= TOL
    assert rel_err(dst2[0].view(shape), np.ascontiguousarray(want_img)) <= TOL


def test_image_as_intensity_prior_path():
    """image_as_intensity=True (model.py:131-140): no seeds, the real image normalised to 0..255 is the
    intensity prior; with every gate off the output is exactly that normalisation."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import bench

    shape = (32, 32, 32)
    gen = bench.build_generator(shape, DEV)
    for st in (gen.spatial_deform, gen.resampled, gen.biasfield, gen.noise, gen.gamma):
        st.prob = 0.0
    rs = np.random.RandomState(0)
    image = torch.from_numpy((rs.rand(*shape) * 900 + 50).astype(np.float32)).to(DEV)
    seg = torch.from_numpy(rs.randint(0, 8, shape).astype(np.float32)).to(DEV)
    np.random.seed(0)
    out, sg, img2, params = gen.sample(image=image, segmentation=seg, seeds=None)
    want = (image - image.min()) / (image.max() - image.min()) * 255
    assert float((out - want).abs().max()) <= 255 * 2e-6
    assert torch.equal(sg, seg) and torch.equal(img2, image)
    assert params["selected_seeds"] == {} and params["seed_intensities"] == {}


# ----------------------------------------------------------------------------- min/max pass of the up-sampling
@pytest.mark.parametrize("shape,coarse,kind", [
    ((64, 64, 128), (49, 50, 99), "noise_zero_background"),
    ((64, 64, 128), (64, 64, 128), "noise_zero_background"),
    ((48, 40, 72), (17, 23, 31), "noise"),
    ((48, 40, 72), (30, 30, 50), "negative"),
    ((48, 40, 72), (30, 30, 50), "flat"),
    ((48, 40, 72), (30, 30, 50), "plateau"),
    ((32, 32, 32), (11, 32, 20), "single_peak_corner"),
])
def test_zoom_minmax_equals_extrema_of_zoomed_volume(shape, coarse, kind):
    """fsg_zoom_minmax must return, bit for bit, the extrema of what fsg_zoom writes (the write pass maps
    the maximum to exactly 1)."""
    from fetalsyngen_b200.engine import engine_for

    eng = engine_for(DEV, shape)
    rs = np.random.RandomState(len(kind) * 7 + coarse[0])
    x = rs.randn(*coarse).astype(np.float32) * 20 + 100
    if kind == "noise_zero_background":
        g = np.meshgrid(*[np.linspace(-1, 1, n) for n in coarse], indexing="ij")
        x = np.where(sum(a * a for a in g) < 0.5, x, 0).astype(np.float32)
        x = np.maximum(x, 0)
    elif kind == "negative":
        x = -x
    elif kind == "flat":
        x[:] = 3.25
    elif kind == "plateau":
        x = np.minimum(x, 110).astype(np.float32)
    elif kind == "single_peak_corner":
        x[:] = 1.0
        x[-1, -1, -1] = 7.0
        x[0, 0, 0] = -2.0
    src = torch.from_numpy(x).to(DEV).contiguous()
    factors = [s / c for s, c in zip(shape, coarse)]
    dst = torch.empty((1, *shape), dtype=torch.float32, device=DEV)
    eng.zoom([src], [coarse], [factors], dst, post=0)
    want = (float(dst.min()), float(dst.max()))
    mm = eng.zoom([src], [coarse], [factors], torch.empty_like(dst), post=1)
    torch.cuda.synchronize()
    got = tuple(float(v) for v in mm[0].cpu())
    assert got == want, (got, want)


@pytest.mark.parametrize("shape,coarses", [((40, 72, 128), [(17, 31, 50), (40, 72, 128)]), ((33, 47, 256), [(12, 16, 86), (30, 40, 201)]), ((24, 20, 512), [(9, 7, 170), (24, 20, 300)]),
                                            ((31, 45, 38), [(11, 15, 13), (31, 45, 38)])])
def test_zoom_kernels_vs_oracle_in_mixed_batches(shape, coarses):
    """myzoom_torch (utils/generation.py:310-397) through both up-sampling kernels — the walk kernel (sz in {128, 256,
    512}: axes blended z, y, x) and the generic plane kernel — on non-cubic volumes, two jobs with different coarse
    grids per launch, against the oracle's x, y, z order (they differ by rounding only)."""
    eng = engine_for(DEV, shape, (0.5, 0.5, 0.5))
    rs = np.random.RandomState(shape[2])
    srcs, want = [], []
    for c in coarses:
        x = (rs.rand(*c) * 300).astype(np.float32)
        srcs.append(torch.from_numpy(x).to(DEV).contiguous())
        want.append(O.zoom_linear(x, np.asarray(shape, dtype=np.float64) / np.asarray(c, dtype=np.float64)))
    dst = torch.empty((len(coarses), int(np.prod(shape))), dtype=torch.float32, device=DEV)
    eng.zoom(srcs, coarses, [[s / c for s, c in zip(shape, cc)] for cc in coarses], dst, post=0)
    for b in range(len(coarses)):
        assert want[b].shape == tuple(shape)
        assert rel_err(dst[b].view(shape), want[b]) <= 1e-6
