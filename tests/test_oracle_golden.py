"""CPU: pin the numpy oracle against the golden vectors made by the unmodified reference."""
import numpy as np
import pytest

import np_oracle as O
from golden_util import BASE_CASES, load_case, oracle_params

TOL = 1e-4  # max-abs relative to the intensity range (north star)


def rel_err(a, b):
    rng = float(b.max() - b.min()) or 1.0
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) / rng


def test_cases_present():
    assert len(BASE_CASES) >= 6


def test_philox_kat():
    # Random123 known-answer vectors for philox4x32-10
    assert O.philox4x32_10([0] * 4, [0] * 2) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert O.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert O.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


@pytest.mark.parametrize("name", BASE_CASES)
def test_gmm_and_means(name):
    d = load_case(name)
    mus0 = (np.float32(25) + np.float32(200) * d["mus_u"]).astype(np.float32)
    mus = O.tie_subclass_means(mus0, d["seed_labels"], d["generation_classes"], d["mus_perturb"])
    np.testing.assert_array_equal(mus, d["mus"])
    sig = (np.float32(5) + np.float32(20) * d["sigmas_u"]).astype(np.float32)
    np.testing.assert_array_equal(sig, d["sigmas"])
    labels = sum(d[f"seed_m{m}"].astype(np.int64) for m in range(1, 5))
    np.testing.assert_array_equal(labels, d["labels"])
    if "intensity" in d:
        out = O.gmm_intensities(d["labels"], d["mus"], d["sigmas"], d["gmm_noise"])
        np.testing.assert_array_equal(out, d["intensity"])


@pytest.mark.parametrize("name", [c for c in BASE_CASES if c not in ("c32_gates_off",)])
def test_affine_field_coords(name):
    d = load_case(name)
    A = O.make_affine_matrix(d["rotations"], d["shears"], d["scalings"]).astype(np.float32)
    np.testing.assert_array_equal(A, d["A"])
    shape = tuple(int(v) for v in d["shape"])
    F = None
    if "Fsmall_n" in d:
        fs = (np.float32(d["nonlin_std"]) * d["Fsmall_n"]).astype(np.float32)
        F = O.zoom_linear(fs, np.asarray(shape) / np.asarray(fs.shape[:3]))
        if "F" in d:
            np.testing.assert_array_equal(F, d["F"])
    if "coords" in d:
        cc = O.deformation_coords(shape, shape, d["A"], d["c2"], F)
        for a in range(3):
            np.testing.assert_array_equal(cc[a], d["coords"][a])


@pytest.mark.parametrize("name", BASE_CASES)
def test_full_pipeline(name):
    d = load_case(name)
    out, seg, st = O.generate_base(d["labels"], d["seg_in"], oracle_params(d))
    # segmentation: bit exact
    np.testing.assert_array_equal(seg, d["seg_out"])
    for key, gkey in (("intensity", "intensity"), ("gamma", "gamma_out"), ("bias", "bias_out"), ("lowres", "lowres"), ("noisy", "noisy")):
        if gkey in d:
            assert st[key].shape == d[gkey].shape
            assert rel_err(st[key], d[gkey]) <= TOL, (key, rel_err(st[key], d[gkey]))
    assert rel_err(out, d["final"]) <= TOL
    assert rel_err(O.scale_intensity(out), d["scaled"]) <= TOL


# ------------------------------------------------------------------ BASELINE.json configs[0] at full size
@pytest.mark.parametrize("name", ["sta30_all_flip", "sta38_noresample"])
def test_full_size_pipeline_vs_reference(name):
    """The oracle at 256^3 on the bundled subjects against the unmodified reference (tests/golden/full_*.npz):
    warped segmentation bit-exact over the whole volume, the image on the stored strided sample, and the
    per-plane sums of the GMM image and of the result."""
    from golden_util import SUB, load_full_case

    d, labels, seg_in, p = load_full_case(name)
    out, seg, st = O.generate_base(labels, seg_in, p)
    np.testing.assert_array_equal(seg, d["seg_out"])
    rng = float(d["final_max"] - d["final_min"])
    assert np.abs(out[SUB].astype(np.float64) - d["final_sub"]).max() / rng <= TOL
    assert float(out.max()) == pytest.approx(float(d["final_max"]), rel=1e-6)
    nplane = out.shape[1] * out.shape[2]
    assert np.abs(out.astype(np.float64).sum(axis=(1, 2)) - d["final_plane_sums"]).max() / nplane / rng <= TOL
    assert np.abs(st["intensity"].astype(np.float64).sum(axis=(1, 2)) - d["intensity_plane_sums"]).max() / nplane <= 1e-3
