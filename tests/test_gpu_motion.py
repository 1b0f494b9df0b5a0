"""SimulateMotion on the GPU: slice acquisition / PSF reconstruction kernels against (a) the
reference's own CUDA extension built from its sources into oracle/_ref, (b) the numpy oracle,
(c) golden vectors of the unmodified reference Scanner + PSFReconstructor.

Float tolerance: max-abs <= 1e-4 x intensity range (TOL).  The reconstruction scatters with
float atomics (summation order is not deterministic, in the reference either) and rounds tap
positions to the nearest voxel, where a 1-ulp coordinate difference can move one tap to the
neighbouring voxel: for it the bound is on the 99.9th percentile, with a loose cap on the max."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import np_motion as M
from gpu_util import DEV, TOL

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
ROOT = Path(__file__).resolve().parents[1]


def load(name):
    with np.load(GOLDEN / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    rng = float(b.max() - b.min()) or 1.0
    return np.abs(a.astype(np.float64).reshape(b.shape) - b.astype(np.float64)) / rng


def close(e, tol=TOL, cap=5e-2, q=99.9, frac=None):
    """Float parity up to rare in/out flips of single PSF taps at the volume faces (a tap whose
    coordinate sits within an ulp of a bound is counted by one implementation and not the other).
    ``frac``: additional bound on the fraction of elements above ``tol``."""
    p = float(np.percentile(e, q))
    assert p <= tol and float(e.max()) <= cap, (p, float(e.max()), float((e > tol).mean()))
    if frac is not None:
        assert float((e > tol).mean()) <= frac, float((e > tol).mean())


def random_case(seed, D=40, n=10, hw=48):
    rs = np.random.RandomState(seed)
    vol = rs.rand(D, D + 4, D - 6).astype(np.float32)
    ax = np.concatenate([rs.randn(n, 3) * 0.6, rs.randn(n, 2) * 3, np.linspace(-15, 15, n)[:, None]], 1).astype(np.float32)
    mat = M.axisangle2mat(ax)
    psf = M.get_psf(res_ratio=(1.3, 1.3, 4.2))
    return vol, mat, psf, (hw, hw + 8), 1.3


def ours_forward(mat, vol, psf, shape, res):
    from fetalsyngen_b200.generator.artifacts.simulate_reco import slice_acquisition

    return slice_acquisition(mat, torch.from_numpy(vol).to(DEV), psf, shape, res)[:, 0]


def ours_adjoint(mat, psf, slices, vshape, res, idx=None):
    from fetalsyngen_b200.generator.artifacts.simulate_reco import slice_acquisition_adjoint

    s = slices if isinstance(slices, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(slices)).to(DEV)
    v, w = slice_acquisition_adjoint(mat, psf, s, vshape, res, slice_idx=idx)
    return v[0, 0], w[0, 0]


@pytest.mark.parametrize("seed", [0, 1])
def test_kernels_vs_numpy_oracle(seed):
    vol, mat, psf, shape, res = random_case(seed)
    want = M.slice_acq_forward(mat, vol, psf, shape, res)
    got = ours_forward(mat, vol, psf, shape, res)
    close(rel(got, want))
    assert ((got.cpu().numpy() == 0) == (want == 0)).mean() > 0.999
    wv, ww = M.slice_acq_adjoint(mat, psf, want, vol.shape, res, True)
    gv, gw = ours_adjoint(mat, psf, want, vol.shape, res)
    close(rel(gv, wv))


def _reference_extension():
    """oracle/_ref/slice_acq_cuda.so (built here by __graft_entry__.build(), it travels with the repo): on a CUDA
    box its absence is a failure of the parity set-up, not a reason to skip."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import build_ref

    ext = build_ref.load_built()
    assert ext is not None, "oracle/_ref/slice_acq_cuda.so is missing: run __graft_entry__.build() in the build container before shipping the tree"
    return ext


@pytest.mark.parametrize("res_s,thick,gap", [(0.6, 2.5, 3.5), (1.0, 3.5, 1.6), (0.3, 1.5, 5.0)])
def test_kernels_vs_reference_extension_at_full_size(res_s, thick, gap):
    """The three 256^3 shapes that profiles/*_motion.jsonl times (288^2 / 160^2 / 544^2 slices; 215 / 729 / 37
    PSF taps), on the bundled sub-sta30 label map: acquisition and PSF reconstruction asserted against the
    reference's own extension, including the xy-quad source copy the production path acquires from."""
    from fetalsyngen_b200.generator.artifacts import simulate_reco as SR
    from fetalsyngen_b200.generator.artifacts import svort
    from golden_util import load_subject

    ext = _reference_extension()
    seg, _, _ = load_subject("sub-sta30")
    S = seg.shape[0]
    rs = np.random.RandomState(5)
    vol = torch.from_numpy((seg > 0).astype(np.float32) * (0.3 + 0.7 * rs.rand(S, S, S).astype(np.float32))).to(DEV)
    np.random.seed(1)
    psf = svort.get_PSF(res_ratio=(res_s / 0.5, res_s / 0.5, thick / 0.5))
    ss = int(np.ceil(int(np.sqrt(3 * S * S / 2.0) * 0.5 / res_s) / 32.0) * 32)
    ns = int(S * 0.5 / gap) + 2
    init = svort.random_init_stack_transforms(ns, gap, False, 3.0)
    motion = svort.sample_motion(np.arange(ns) * 1.5, True)
    mat = svort.mat_update_resolution(motion.compose(init).matrix(), 0.5, 0.5)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    empty = torch.empty(0, device=DEV)
    ref_s = ext.forward(t(mat), vol[None, None], empty, empty, t(psf), [ss, ss], float(res_s / 0.5), False, False)[0]
    ref_np = ref_s[:, 0].cpu().numpy()
    for pairs in (None, SR.volume_xyquads(vol)):
        got = SR.slice_acquisition(mat, vol, psf, (ss, ss), res_s / 0.5, pairs=pairs)[:, 0]
        close(rel(got, ref_np))
    nstack = max(1, min(6, 250 // ns))
    mats = np.concatenate([mat] * nstack)[:250]
    sl = torch.cat([ref_s] * nstack)[:250].contiguous()
    ref_v = ext.adjoint_forward(t(mats), t(psf), sl, empty, empty, [S, S, S], float(res_s / 0.5), True, True)[0][0, 0].cpu().numpy()
    gv, _ = SR.slice_acquisition_adjoint(mats, psf, sl, (S, S, S), res_s / 0.5)
    # the reconstruction normalises by the accumulated PSF weight: on this sparse (brain-only) volume a voxel
    # reached by one or two taps changes by a sizeable part of the range when a single tap rounds to the
    # neighbouring voxel in one implementation (r02: 1.5e-4 of the voxels above TOL, max 0.07), so the cap on
    # the maximum is loose and the fraction of such voxels is bounded instead
    close(rel(gv[0, 0], ref_v), cap=0.25, frac=5e-4)


@pytest.mark.parametrize("seed", [0, 1])
def test_kernels_and_oracle_vs_reference_extension(seed):
    """The reference's slice_acq_cuda extension (built by oracle/build_ref.py from /root/reference)
    is the ground truth for both our kernels and the numpy restatement."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import build_ref

    ext = _reference_extension()
    vol, mat, psf, shape, res = random_case(seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    empty = torch.empty(0, device=DEV)
    ref_s = ext.forward(t(mat), t(vol)[None, None], empty, empty, t(psf), list(shape), float(res), False, False)[0][:, 0]
    ref_np = ref_s.cpu().numpy()
    close(rel(M.slice_acq_forward(mat, vol, psf, shape, res), ref_np))
    close(rel(ours_forward(mat, vol, psf, shape, res), ref_np))
    ref_v = ext.adjoint_forward(t(mat), t(psf), ref_s[:, None].contiguous(), empty, empty, list(vol.shape), float(res), True, True)[0][0, 0].cpu().numpy()
    for got in (M.slice_acq_adjoint(mat, psf, ref_np, vol.shape, res, True)[0], ours_adjoint(mat, psf, ref_np, vol.shape, res)[0]):
        close(rel(got, ref_v))


@pytest.mark.parametrize("name", ["motion_default", "motion_all_on"])
def test_forward_and_adjoint_vs_reference_golden(name):
    g = load(name)
    for k in range(int(g["n_attempts"])):
        got = ours_forward(g[f"fwd_mat_{k}"], g["image"], g["psf_acq"], g[f"fwd_img_{k}"].shape[1:], float(g["resolution_slice"]) / 0.5)
        close(rel(got, g[f"fwd_img_{k}"]))
    mask = (g["seg"] > 0).astype(np.float32)
    got = ours_forward(g["fwd_mat_0"], mask, M.get_psf(0), g["fwd_mask_0"].shape[1:], float(g["resolution_slice"]) / 0.5)
    close(rel(got, g["fwd_mask_0"]))
    n = g["stacks"].shape[0]
    kept = g["perm_kept"][int(n * float(g["rm_slices_ratio"])):] if bool(g["rm_slices_on"]) else np.arange(n)
    assert len(kept) == int(g["adj_nslices"])
    gv, _ = ours_adjoint(g["adj_mat"], g["psf_rec"], g["stacks"], g["image"].shape, float(g["res_slice"]), idx=kept)
    close(rel(gv, g["recon"]))


def _inject(g):
    inj = {}
    h, w = g["stacks"].shape[1:]
    for k in range(int(g["n_stacks_logged"])):
        n = g[f"noise_mask_{k}"].size * 8 // (h * w)
        mask = np.unpackbits(g[f"noise_mask_{k}"])[: n * h * w].reshape(n, h, w).astype(bool)
        for key in ("noise1", "noise2"):
            full = np.zeros(mask.shape, np.float32)
            full[mask] = g[f"{key}_{k}"]
            inj[f"{key}_{k}"] = full
        inj[f"void_{k}"] = {kk: g[f"void_{kk}_{k}"] for kk in ("idx", "yc", "xc", "theta", "a", "A", "sx") if f"void_{kk}_{k}" in g}
    for key in ("perm_misreg", "perm_kept"):
        if key in g:
            inj[key] = g[key]
    for o in range(int(g["octave"])):
        inj[f"theta_{o}"], inj[f"phi_{o}"] = g[f"theta_{o}"], g[f"phi_{o}"]
    return inj


def _artifact(scanner_over=None, recon_over=None):
    from fetalsyngen_b200.generator.artifacts.utils import ReconMergeParams, ReconParams, ScannerParams
    from fetalsyngen_b200.generator.augmentation.artifacts import SimulateMotion

    sp = dict(resolution_slice_fac_min=0.5, resolution_slice_fac_max=2, resolution_slice_max=1.5, slice_thickness_min=1.5, slice_thickness_max=3.5, gap_min=1.5, gap_max=5.5,
              min_num_stack=2, max_num_stack=6, max_num_slices=250, noise_sigma_min=0, noise_sigma_max=0.1, TR_min=1, TR_max=2, prob_void=0.2, prob_gamma=0.1, gamma_std=0.05,
              slice_size=None, restrict_transform=False, txy=3.0)
    sp.update(scanner_over or {})
    mp = ReconMergeParams(merge_type="perlin", perlin_res_list=[1, 2], perlin_octaves_list=[1, 2, 4], perlin_persistence=0.5, perlin_lacunarity=2,
                          gauss_ngaussians_min=2, gauss_ngaussians_max=4, perlin_increase_size=0.25)
    rp = dict(prob_misreg_slice=0.1, slices_misreg_ratio=0.1, prob_misreg_stack=0.1, txy=3.0, prob_merge=1.0, merge_params=mp, prob_smooth=0.2, prob_rm_slices=0.3,
              rm_slices_min=0.1, rm_slices_max=0.4)
    rp.update(recon_over or {})
    return SimulateMotion(1.0, ScannerParams(**sp), ReconParams(**rp))


CASES = {
    "motion_default": ({"max_num_stack": 3}, None),
    "motion_all_on": ({"prob_gamma": 1.0, "prob_void": 0.5, "min_num_stack": 3, "max_num_stack": 3},
                      {"prob_misreg_slice": 1.0, "prob_misreg_stack": 1.0, "prob_smooth": 1.0, "prob_rm_slices": 1.0}),
}


@pytest.mark.parametrize("name", list(CASES))
def test_simulate_motion_vs_reference_golden(name):
    """Whole artifact with the reference's numpy seed (same scalar draws in the same order) and
    its torch draws injected: stacks, transforms and the merged output must match."""
    from fetalsyngen_b200.generator.artifacts import simulate_reco as SR

    g = load(name)
    art = _artifact(*CASES[name])
    img, seg = torch.from_numpy(g["image"]).to(DEV), torch.from_numpy(g["seg"].astype(np.float32)).to(DEV)
    cap = {}
    o_scan = SR.Scanner.scan

    def scan(self, data, genparams={}, inject=None):
        out = o_scan(self, data, genparams, inject)
        cap.update({k: out[k] for k in ("stacks", "transforms", "transforms_gt", "positions")})
        return out

    SR.Scanner.scan = scan
    try:
        np.random.seed(int(g["seed"]))  # the golden run starts at Scanner.scan (after the gate draw of artifacts.py:388)
        out, meta = SR.simulate_motion(art, img, seg, [0.5, 0.5, 0.5], inject=_inject(g))
    finally:
        SR.Scanner.scan = o_scan
    assert cap["stacks"].shape[0] == g["stacks"].shape[0]
    assert np.array_equal(cap["positions"], g["positions"])
    assert np.abs(cap["transforms_gt"] - g["transforms_gt"]).max() <= 1e-4
    assert np.abs(cap["transforms"] - g["transforms"]).max() <= 1e-4
    close(rel(cap["stacks"][:, 0], g["stacks"]), 2 * TOL)
    assert meta["nstacks"] == len(np.unique(g["positions"][:, 1]))
    assert bool(meta["smooth_volume_on"]) == bool(g["smooth_volume_on"]) and int(meta["octave"]) == int(g["octave"])
    close(rel(out, g["output"]), 2 * TOL)


def test_simulate_motion_philox_mode_and_gate():
    g = load("motion_default")
    img, seg = torch.from_numpy(g["image"]).to(DEV), torch.from_numpy(g["seg"].astype(np.float32)).to(DEV)
    art = _artifact()
    art.prob = 0.0
    out, meta = art(img, seg, DEV, {}, resolution=[0.5, 0.5, 0.5])
    assert out is img and meta == {}
    art.prob = 1.0
    np.random.seed(5)
    torch.manual_seed(5)
    out, meta = art(img, seg, DEV, {}, resolution=[0.5, 0.5, 0.5])
    assert out.shape == img.shape and torch.isfinite(out).all()
    for key in ("resolution_recon", "resolution_slice", "slice_thickness", "gap", "nstacks", "smooth_volume_on", "rm_slices_on", "rm_slices_ratio", "misreg_stack_on", "misreg_slice_on",
                "merge_volume_on", "merge_type", "res", "octave"):
        assert key in meta
    diff = (out - img).abs()
    assert float(diff.max()) > 1e-3            # something was degraded ...
    assert float(out.min()) >= -1e-6 and float(out.max()) <= float(img.max()) * 1.5 + 0.5


@pytest.mark.parametrize("psf_ratio", [(1.2, 1.2, 5.0), (2.0, 2.0, 7.0)])
def test_xpairs_acquisition_is_bit_identical(psf_ratio):
    """fsg_slice_acq_forward_xpairs (two x corners per 8-byte load from the pair volume) must give exactly
    the slices of fsg_slice_acq_forward, for the thread-per-pixel and the lanes-over-taps kernels."""
    from fetalsyngen_b200.generator.artifacts import simulate_reco as SR
    from fetalsyngen_b200.generator.artifacts import svort

    rs = np.random.RandomState(4)
    vol = torch.from_numpy(rs.rand(40, 44, 48).astype(np.float32)).to("cuda:0")
    psf = svort.get_PSF(res_ratio=psf_ratio)
    ax = np.concatenate([rs.randn(9, 3) * 0.5, rs.randn(9, 2) * 3, np.linspace(-15, 15, 9)[:, None]], 1).astype(np.float32)
    mat = svort.axisangle2mat(ax)
    a = SR.slice_acquisition(mat, vol, psf, (64, 64), 1.2)
    b = SR.slice_acquisition(mat, vol, psf, (64, 64), 1.2, pairs=SR.volume_xpairs(vol))
    assert torch.equal(a, b) and float(a.abs().max()) > 0


def test_xyquads_acquisition_is_bit_identical():
    from fetalsyngen_b200.generator.artifacts import simulate_reco as SR
    from fetalsyngen_b200.generator.artifacts import svort

    rs = np.random.RandomState(6)
    vol = torch.from_numpy(rs.rand(40, 44, 48).astype(np.float32)).to("cuda:0")
    psf = svort.get_PSF(res_ratio=(1.2, 1.2, 5.0))
    ax = np.concatenate([rs.randn(9, 3) * 0.5, rs.randn(9, 2) * 3, np.linspace(-15, 15, 9)[:, None]], 1).astype(np.float32)
    mat = svort.axisangle2mat(ax)
    a = SR.slice_acquisition(mat, vol, psf, (64, 64), 1.2)
    b = SR.slice_acquisition(mat, vol, psf, (64, 64), 1.2, pairs=SR.volume_xyquads(vol))
    assert torch.equal(a, b) and float(a.abs().max()) > 0


# ------------------------------------------------------------------ the pybind modules' full contract (native_compat)
def _compat_case(seed):
    vol, mat, psf, shape, res = random_case(seed, D=36, n=6, hw=40)
    rs = np.random.RandomState(100 + seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    vol_mask = rs.rand(*vol.shape) > 0.2
    slices_mask = rs.rand(mat.shape[0], 1, *shape) > 0.3
    return t(vol)[None, None].contiguous(), t(mat), t(psf), shape, float(res), t(vol_mask)[None, None].contiguous(), t(slices_mask).contiguous()


@pytest.mark.parametrize("interp_psf", [False, True])
@pytest.mark.parametrize("need_weight", [False, True])
@pytest.mark.parametrize("masks", ["none", "vol", "slices", "both"])
def test_forward_full_contract_vs_reference_extension(interp_psf, need_weight, masks):
    """slice_acq_cuda.forward with every option (volume / slice masks, weight output, both PSF modes): the drop-in
    module over libfsg against the reference's own extension, same argument list."""
    from fetalsyngen_b200.generator.artifacts.native_compat import slice_acq_cuda as ours

    ext = _reference_extension()
    vol, mat, psf, shape, res, vmask, smask = _compat_case(3)
    empty_b = torch.empty(0, dtype=torch.bool, device=DEV)
    vm = vmask if masks in ("vol", "both") else empty_b
    sm = smask if masks in ("slices", "both") else empty_b
    want = ext.forward(mat, vol, vm, sm, psf, list(shape), res, need_weight, interp_psf)
    got = ours.forward(mat, vol, vm, sm, psf, list(shape), res, need_weight, interp_psf)
    assert len(got) == len(want) == (2 if need_weight else 1)
    for g, w in zip(got, want):
        assert g.shape == w.shape
        close(rel(g, w.cpu().numpy()))
        assert ((g == 0) == (w == 0)).float().mean() > 0.999  # the same pixels stay unwritten
    if masks in ("slices", "both"):
        assert float(got[0][~smask].abs().max()) == 0.0


@pytest.mark.parametrize("interp_psf", [False, True])
@pytest.mark.parametrize("equalize", [False, True])
@pytest.mark.parametrize("masks", ["none", "vol", "slices", "both"])
def test_adjoint_full_contract_vs_reference_extension(interp_psf, equalize, masks):
    from fetalsyngen_b200.generator.artifacts.native_compat import slice_acq_cuda as ours

    ext = _reference_extension()
    vol, mat, psf, shape, res, vmask, smask = _compat_case(4)
    empty_b, empty_f = torch.empty(0, dtype=torch.bool, device=DEV), torch.empty(0, device=DEV)
    slices = ext.forward(mat, vol, empty_f, empty_f, psf, list(shape), res, False, False)[0]
    vm = vmask if masks in ("vol", "both") else empty_b
    sm = smask if masks in ("slices", "both") else empty_b
    want = ext.adjoint_forward(mat, psf, slices, sm, vm, list(vol.shape[-3:]), res, interp_psf, equalize)
    got = ours.adjoint_forward(mat, psf, slices, sm, vm, list(vol.shape[-3:]), res, interp_psf, equalize)
    close(rel(got[0], want[0].cpu().numpy()), cap=0.25, frac=2e-3)
    if equalize:
        close(rel(got[1], want[1].cpu().numpy()), cap=0.25, frac=2e-3)
    if masks in ("vol", "both"):
        assert float(got[0][~vmask].abs().max()) == 0.0


def test_transform_conversions_vs_reference_extension_and_oracle():
    """transform_convert_cuda.axisangle2mat_forward / mat2axisangle_forward: small and large angles, the
    first-order branch below 1e-6, all four quaternion branches, against the reference's module and the oracle."""
    from fetalsyngen_b200.generator.artifacts.native_compat import transform_convert_cuda as ours

    sys.path.insert(0, str(ROOT / "oracle"))
    import build_ref

    rs = np.random.RandomState(8)
    ax = np.concatenate([rs.randn(200, 3) * 1.5, rs.randn(200, 3) * 20], 1).astype(np.float32)
    ax[:10, :3] *= 1e-4                      # first-order branch
    ax[10:40, :3] *= 3.14159 / np.linalg.norm(ax[10:40, :3], axis=1, keepdims=True)  # rotations by ~pi: the three trace-negative branches
    axd = torch.from_numpy(ax).to(DEV)
    mat = ours.axisangle2mat_forward(axd)[0]
    assert np.abs(mat.cpu().numpy() - M.axisangle2mat(ax)).max() <= 2e-6
    back = ours.mat2axisangle_forward(mat)[0]
    assert np.abs(back.cpu().numpy() - M.mat2axisangle(mat.cpu().numpy())).max() <= 2e-4  # atan2 / sqrt near pi amplify an ulp of the matrix
    mat2 = ours.axisangle2mat_forward(back)[0]
    assert float((mat2 - mat).abs().max()) <= 5e-4  # same rotation after the round trip
    ext = build_ref.load_built("transform_convert_cuda")
    assert ext is not None, "oracle/_ref/tc/transform_convert_cuda.so is missing: run __graft_entry__.build() in the build container"
    assert float((ext.axisangle2mat_forward(axd)[0] - mat).abs().max()) <= 2e-6
    assert float((ext.mat2axisangle_forward(mat)[0] - back).abs().max()) <= 2e-4
