"""CPU tests of the host side of SimulateMotion (transforms, PSF, trajectories, draw order)
against the numpy oracle, scipy and the golden vectors of the unmodified reference."""
from pathlib import Path

import numpy as np
import pytest

import np_motion as M
from fetalsyngen_b200.generator.artifacts import svort

GOLDEN = Path(__file__).resolve().parent / "golden"


def load(name):
    with np.load(GOLDEN / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


def test_transform_conversions_match_oracle():
    rs = np.random.RandomState(3)
    ax = np.concatenate([rs.randn(64, 3) * 1.2, rs.randn(64, 3) * 5], 1).astype(np.float32)
    ax[:4, :3] *= 1e-4  # small-angle branch
    assert np.abs(svort.axisangle2mat(ax) - M.axisangle2mat(ax)).max() <= 1e-6
    mat = M.axisangle2mat(ax)
    assert np.abs(svort.mat2axisangle(mat) - M.mat2axisangle(mat)).max() <= 1e-5
    # all four quaternion branches: rotations by ~pi about each axis
    for axis in np.eye(3):
        a = np.concatenate([axis * 3.1, [1, 2, 3]])[None].astype(np.float32)
        m = M.axisangle2mat(a)
        assert np.abs(svort.mat2axisangle(m) - M.mat2axisangle(m)).max() <= 1e-5
        assert np.abs(M.axisangle2mat(M.mat2axisangle(m)) - m).max() <= 1e-5


@pytest.mark.parametrize("name", ["motion_default", "motion_all_on"])
def test_psf_matches_reference_golden(name):
    g = load(name)
    rs, th = float(g["resolution_slice"]), float(g["slice_thickness"])
    for fn in (svort.get_PSF, M.get_psf):
        psf = fn(res_ratio=(rs / 0.5, rs / 0.5, th / 0.5))
        assert psf.shape == g["psf_acq"].shape
        assert np.abs(psf - g["psf_acq"]).max() <= 1e-7
    taps, radius = svort.psf_taps(g["psf_acq"])
    assert np.array_equal(taps, M.psf_taps(g["psf_acq"]))
    assert radius >= np.sqrt((taps[:, :3] ** 2).sum(1)).max()
    assert svort.get_PSF(0).shape == (1, 1, 1) and svort.get_PSF(0)[0, 0, 0] == 1


def test_rigid_transform_algebra():
    rs = np.random.RandomState(0)
    ax = np.concatenate([rs.randn(8, 3), rs.randn(8, 3) * 4], 1).astype(np.float32)
    t = svort.RigidTransform(ax, trans_first=True)
    assert np.abs(t.axisangle() - ax).max() == 0
    # trans_first <-> trans_last round trip
    m_last = t.matrix(trans_first=False)
    back = svort.RigidTransform(m_last, trans_first=False).matrix(trans_first=True)
    assert np.abs(back - t.matrix()).max() <= 1e-5
    # compose with the identity and associativity of the rotation part
    ident = svort.RigidTransform(np.zeros((8, 6), np.float32))
    assert np.abs(t.compose(ident).matrix() - t.matrix()).max() <= 1e-6
    u = svort.RigidTransform(ax[::-1].copy())
    c = t.compose(u).matrix()
    assert np.abs(c[:, :, :3] - np.einsum("nij,njk->nik", t.matrix()[:, :, :3], u.matrix()[:, :, :3])).max() <= 1e-6
    assert len(t[2:5]) == 3 and len(t[3]) == 1
    r = svort.reset_transform(t)
    assert np.abs(r.axisangle()[:, :5]).max() == 0 and abs(float(r.axisangle()[:, 5].mean())) < 1e-5


def test_interleave_and_trajectories():
    assert svort.interleave_index(7, 3) == M.interleave_index(7, 3) == [0, 3, 5, 1, 4, 6, 2]
    from scipy.interpolate import interp1d

    tr = svort.get_trajectory()
    assert len(tr["rot_T"]) == 154 and len(tr["trans_T"]) == 154
    y = tr["rot_y"][tr["rot_off"][3] : tr["rot_off"][4]]
    f = interp1d(np.arange(len(y)), y, axis=0, fill_value="extrapolate")
    t = np.array([0.0, 0.4, 1.0, 17.25, len(y) - 1.0, len(y) + 2.5])
    assert np.abs(svort._traj_eval(y, t) - f(t)).max() <= 1e-12
    np.random.seed(0)
    m = svort.sample_motion(np.arange(12) * 1.5)
    mat = m.matrix(trans_first=False)
    assert np.abs(mat[0] - np.eye(3, 4)).max() <= 1e-6  # motion is relative to the first slice
    assert np.abs(np.linalg.det(mat[:, :, :3].astype(np.float64)) - 1).max() <= 1e-5


@pytest.mark.parametrize("name,over", [("motion_default", {"max_num_stack": 3}), ("motion_all_on", {"min_num_stack": 3, "max_num_stack": 3})])
def test_host_draw_order_reproduces_reference_transforms(name, over):
    """With numpy seeded like the golden run, the host code up to the first acquisition must draw the
    same scalars in the same order as the reference: same resolution / thickness / gap and the
    same slice transforms for the first stack."""
    from fetalsyngen_b200.generator.artifacts.simulate_reco import Scanner

    g = load(name)
    sp = dict(resolution_slice_fac_min=0.5, resolution_slice_fac_max=2, resolution_slice_max=1.5, slice_thickness_min=1.5, slice_thickness_max=3.5, gap_min=1.5, gap_max=5.5,
              min_num_stack=2, max_num_stack=6, max_num_slices=250, noise_sigma_min=0, noise_sigma_max=0.1, TR_min=1, TR_max=2, prob_void=0.2, prob_gamma=0.1, gamma_std=0.05,
              slice_size=None, restrict_transform=False, txy=3.0, resolution_recon=np.float64(0.5))
    sp.update(over)
    sc = Scanner(**sp)
    np.random.seed(int(g["seed"]))
    d = sc.get_resolution({"resolution": np.float64(0.5)})
    assert d["resolution_slice"] == float(g["resolution_slice"]) and d["slice_thickness"] == float(g["slice_thickness"]) and d["gap"] == float(g["gap"])
    ns = int(48 * 0.5 / d["gap"]) + 2
    np.random.randint(sc.min_num_stack, sc.max_num_stack + 1)
    init = svort.random_init_stack_transforms(ns, d["gap"], sc.restrict_transform, sc.txy)
    motion = svort.sample_motion(sc.sample_time(ns), True)
    motion = motion[svort.interleave_index(ns, np.random.randint(2, int(np.sqrt(ns)) + 1))]
    mat = svort.mat_update_resolution(motion.compose(init).matrix(), 0.5, 0.5)
    assert mat.shape == g["fwd_mat_0"].shape
    assert np.abs(mat - g["fwd_mat_0"]).max() <= 1e-4
