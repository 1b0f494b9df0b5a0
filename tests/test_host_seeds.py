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


# ----------------------------------------------------------------------------- bit-packed seed cache (host side)
def _phantom_seeds(shape, smax):
    from fetalsyngen_b200.utils.phantom import label_phantom

    seeds, seg = {}, None
    for n in range(1, smax + 1):
        seg, sv = label_phantom(shape, n_sub=(n, n, n, n), seed=n)
        seeds[n] = {m + 1: sv[m] for m in range(4)}
    return seg, seeds


@pytest.mark.parametrize("smax,dtype", [(6, np.uint16), (10, np.uint32), (1, np.uint16)])
def test_packed_seeds_round_trip(smax, dtype):
    from fetalsyngen_b200.data import packed as K

    _, seeds = _phantom_seeds((20, 24, 28), smax)
    words, counts = K.pack_seed_volumes(seeds)
    assert words.dtype == dtype and counts == list(range(1, smax + 1))
    rs = np.random.RandomState(0)
    for _ in range(8):
        m2s = {m: int(rs.randint(1, smax + 1)) for m in range(1, 5)}
        want = sum(seeds[m2s[m]][m].astype(np.int32) for m in range(1, 5)).astype(np.uint8)  # rand_gmm.py:90-97
        assert np.array_equal(K.unpack_numpy(words, counts, m2s), want)


def test_packed_seeds_reject_what_cannot_be_packed():
    from fetalsyngen_b200.data import packed as K

    _, seeds = _phantom_seeds((12, 12, 12), 3)
    bad = {n: {m: v.copy() for m, v in per.items()} for n, per in seeds.items()}
    bad[2][1][bad[2][2] != 0] = 10  # overlapping supports
    with pytest.raises(ValueError, match="overlap"):
        K.pack_seed_volumes(bad)
    bad = {n: {m: v.copy() for m, v in per.items()} for n, per in seeds.items()}
    bad[3][1][:] = 0  # support changes with the sub-class count
    with pytest.raises(ValueError, match="differs"):
        K.pack_seed_volumes(bad)
    bad = {n: {m: v.copy() for m, v in per.items()} for n, per in seeds.items()}
    bad[2][3][bad[2][3] != 0] = 35  # sub-class index 5 of a 2-class split
    with pytest.raises(ValueError, match="outside"):
        K.pack_seed_volumes(bad)
    with pytest.raises(ValueError):
        K.field_layout([1, 2, 40])


def test_packed_cache_file_and_converter(tmp_path):
    from fetalsyngen_b200.data import packed as K
    from fetalsyngen_b200.utils.nifti import write_nifti

    shape = (16, 20, 24)
    seg, seeds = _phantom_seeds(shape, 4)
    aff = np.diag([0.5, 0.5, 0.5, 1.0])
    segf = tmp_path / "sub-a_dseg.nii.gz"
    write_nifti(segf, seg.astype(np.float32), aff)
    paths = {}
    for n, per in seeds.items():
        paths[n] = {}
        for m, v in per.items():
            paths[n][m] = tmp_path / f"s{n}_m{m}.nii.gz"
            write_nifti(paths[n][m], v, aff)
    f = K.pack_subject(segf, paths, tmp_path / "cache" / "sub-a.fsgpack.npz")
    seg2, ps, aff2 = K.load_packed(f)
    assert np.array_equal(seg2, seg) and seg2.dtype == np.uint8 and np.allclose(aff2, aff)
    assert ps.counts == [1, 2, 3, 4] and ps.shape == shape and ps.word_bytes == 2
    m2s = {1: 4, 2: 1, 3: 3, 4: 2}
    want = sum(seeds[m2s[m]][m].astype(np.int32) for m in range(1, 5)).astype(np.uint8)
    assert np.array_equal(K.unpack_numpy(ps._host, ps.counts, m2s), want)
    with pytest.raises(KeyError):
        ps.job(type("J", (), {"shift": [0] * 4, "mask": [0] * 4})(), {1: 5, 2: 1, 3: 1, 4: 1}, "cpu", None)
