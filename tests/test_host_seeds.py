"""CPU tests of seed generation: the numpy oracle against golden vectors of
the unmodified scikit-learn GaussianMixture / kmeans_plusplus and of the reference's own
``split_lables`` (tests/golden/make_golden_seeds.py), and the host-side label handling of the product."""
from pathlib import Path

import numpy as np
import pytest

import np_seeds as S
from fetalsyngen_b200 import seeds as P

GOLDEN = Path(__file__).resolve().parent / "golden"
EM_CASES = ["seeds_em_k3", "seeds_em_k6", "seeds_em_k2_small", "seeds_em_k10_overlap"]


def load(name):
    with np.load(GOLDEN / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", EM_CASES)
def test_oracle_em_matches_sklearn_golden(name):
    g = load(name)
    f = S.fit_from_indices(g["x"], g["indices"])
    assert f["n_iter"] == int(g["n_iter"]) and f["converged"] == bool(g["converged"])
    assert abs(f["lower_bound"] - float(g["lower_bound"])) <= 1e-12
    for key in ("weights", "means", "covariances"):
        assert np.abs(f[key] - g[key]).max() <= 1e-10 * max(1.0, np.abs(g[key]).max()), key
    assert np.array_equal(S.predict(g["x"], f), g["labels"])


def test_oracle_split_matches_reference_golden():
    """Whole chain (label fusion, NaN handling, k-means++ with numpy's global stream, best of five EM
    runs, label volumes) against the reference's split_lables, voxel for voxel."""
    g = load("seeds_split")
    for sub in (1, 3):
        out = S.split_labels(g["image"], g["seg"], sub, np.random.RandomState(77))
        for m in range(1, 5):
            assert np.array_equal(out[m], g[f"sub{sub}_m{m}"]), (sub, m)


def test_oracle_kmeans_plusplus_matches_sklearn():
    cluster = pytest.importorskip("sklearn.cluster")
    x = load("seeds_em_k6")["x"]
    for seed in range(6):
        for k in (2, 3, 6, 10):
            _, ref = cluster.kmeans_plusplus(x[:, None].astype(np.float64), k, random_state=np.random.RandomState(seed))
            assert np.array_equal(S.kmeans_plusplus(x, k, np.random.RandomState(seed)), ref)


def test_oracle_rejects_too_few_samples():
    with pytest.raises(ValueError, match="n_samples >= n_components"):
        S.fit_predict(np.array([1.0, 2.0], dtype=np.float32), 3, np.random.RandomState(0))


# ----------------------------------------------------------------------------- product host logic
def test_label_tables_match_reference_maps():
    for ann, table in (("feta", S.FETA2META), ("dhcp", S.DHCP2META)):
        lut = np.frombuffer(P.label_lut(ann), dtype=np.uint8)
        assert lut.size == 256 and lut[0] == 4
        for lab in range(1, 256):
            want = table.get(lab, 0)
            if ann == "dhcp" and lab == 4:
                want = 4  # skull label is background for dHCP (generate_seeds.py:144-146)
            assert lut[lab] == want, (ann, lab)
    with pytest.raises(ValueError):
        P.label_lut("other")
    assert P.FETA2META == S.FETA2META and P.DHCP2META == S.DHCP2META


def test_labels_to_u8_codes():
    seg = np.array([0.0, 1.0, 7.0, np.nan, 2.5, -1.0, 300.0, 254.0, np.inf], dtype=np.float32)
    assert P.labels_to_u8(seg).tolist() == [0, 1, 7, 0, 255, 255, 255, 254, 255]
    u8 = np.arange(6, dtype=np.uint8)
    assert P.labels_to_u8(u8) is not None and np.array_equal(P.labels_to_u8(u8), u8)


def test_meta_labels_oracle_vs_lut():
    """The product's (lut, uint8 code) formulation gives the oracle's meta-label volume."""
    g = load("seeds_split")
    image, meta = S.meta_labels(g["image"], g["seg"])
    lut = np.frombuffer(P.label_lut("feta"), dtype=np.uint8)
    v = lut[P.labels_to_u8(g["seg"])]
    img0 = np.nan_to_num(g["image"], nan=0.0)
    got = np.where(v == 4, np.where(img0 != 0, 4, 0), v).astype(np.uint8)
    assert np.array_equal(got, meta)
    assert np.array_equal(image, img0)


def test_seed_generator_needs_cuda():
    import torch

    from fetalsyngen_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.FsgError):
        P.SeedGenerator()
