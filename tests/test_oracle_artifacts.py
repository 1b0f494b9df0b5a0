"""CPU: pin the numpy artifact oracle against golden vectors made by the unmodified reference."""
import numpy as np
import pytest

import np_artifacts as A
from golden_util import GOLDEN, load_case

TOL = 1e-4


def art(name):
    with np.load(GOLDEN / f"art_{name}.npz", allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def base():
    d = load_case("c64_default")
    return d["final"].astype(np.float32), d["seg_out"]


def rel(a, b):
    rng = float(b.max() - b.min()) or 1.0
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) / rng


def test_blur_cortex():
    g = art("blur_cortex")
    img, seg = base()
    assert rel(A.cortex_prior(img.shape), g["prior"]) <= 1e-6
    # the drawn centres are the multinomial picks among the cortex voxels (C order)
    cortex = np.nonzero(seg == 2)
    picks = np.stack([c[g["multinomial_idx"]] for c in cortex], 1)
    np.testing.assert_array_equal(picks, g["centers"])
    mog = A.mog_3d(img.shape, g["centers"], g["sigmas"])
    assert np.abs(mog - g["gaussian"]).max() <= 1e-5
    assert rel(A.blur_cortex(img, g["centers"], g["sigmas"], g["std_blurs"]), g["output"]) <= TOL


@pytest.mark.parametrize("name", ["struct_noise", "struct_noise_oct"])
def test_struct_noise(name):
    g = art(name)
    img, seg = base()
    n = int(g["nstages"])
    lr = A.multiscale_noise(img.shape, [g[f"randn_{k}"] for k in range(n)])
    if "lr_noise" in g:
        ref_lr = g["lr_noise"] / np.abs(g["lr_noise"]).max()
        assert np.abs(lr - ref_lr).max() <= 1e-5
    oc, res = int(g["octave"]), int(g["res"])
    th, ph = [g[f"theta_{o}"] for o in range(oc)], [g[f"phi_{o}"] for o in range(oc)]
    if "perlin_0" in g:
        assert np.abs(A.perlin_3d(img.shape, (res,) * 3, th[0], ph[0]) - g["perlin_0"]).max() <= 1e-5
    w = A.fractal_noise_3d(img.shape, (res,) * 3, th, ph, 0.5, 2, 0.1)
    assert np.abs(w - g["weight"]).max() <= 1e-4
    out = A.struct_noise(img, seg, lr, float(g["noise_std"]), w)
    assert rel(out, g["output"]) <= TOL


def test_boundaries():
    g = art("boundaries")
    img, seg = base()
    mask = (seg > 0).astype(np.uint8)
    halo = A.build_halo(mask, int(g["halo_radius"]))
    np.testing.assert_array_equal(halo, g["mask_halo"])
    m = halo
    for i in range(int(g["n_generate_fuzzy"])):
        m = A.fuzzy_iteration(m, g[f"perm_{i}"])
        np.testing.assert_array_equal(m, g[f"fuzzy_{i}"])
    surf = np.nonzero((m.astype(np.int32) - halo.astype(np.int32)) > 0)
    picks = np.stack([s[g["perm_centers"][: int(g["n_centers"])]] for s in surf], 1)
    np.testing.assert_array_equal(picks, g["centers"])
    mog = A.mog_3d(img.shape, g["centers"], list(g["sigmas"]))
    assert np.abs(mog - g["mog"]).max() <= 1e-5
    final, idx = A.boundaries_mask(halo, m, g["mog"], int(g["n_generate_fuzzy"]))
    np.testing.assert_array_equal(img * final, g["output"])
