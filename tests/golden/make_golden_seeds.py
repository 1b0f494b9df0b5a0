"""Golden vectors for seed generation, produced in the build container:

    python tests/golden/make_golden_seeds.py

(a) ``seeds_em_*.npz``: the UNMODIFIED scikit-learn ``GaussianMixture`` (the third-party dependency
    that /root/reference/scripts/generate_seeds.py:179-181 calls; pinned 1.6.1 there, 1.9.0 in this
    image) run from injected k-means++ indices: only ``_initialize_parameters`` is overridden, to place
    the one-hot responsibilities of sklearn's own k-means++ branch at recorded sample indices;
    ``_initialize``, the E and M steps, the convergence test and the final predict are sklearn's code.
(b) ``seeds_split.npz``: the UNMODIFIED reference functions ``split_lables`` / ``subsplit_label``
    (generate_seeds.py:175-211, executed from the read-only tree with the monai stand-in of
    ``oracle/ref_import.py``) on a small synthetic subject with ``np.random.seed`` fixed, i.e. the
    whole chain label fusion -> k-means++ -> best of five EM runs -> seed volumes.
"""
from __future__ import annotations

import sys
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
OUT = Path(__file__).resolve().parent


def mixture(n, comps, seed):
    """n float32 samples from a 1-D mixture [(weight, mean, std), ...], clipped at 0 like an MR magnitude image."""
    rs = np.random.RandomState(seed)
    w = np.array([c[0] for c in comps], dtype=np.float64)
    which = rs.choice(len(comps), size=n, p=w / w.sum())
    mu = np.array([c[1] for c in comps])[which]
    sd = np.array([c[2] for c in comps])[which]
    return np.maximum(mu + sd * rs.randn(n), 0).astype(np.float32)


CASES = {
    # name: (n, components, k, seed)
    "seeds_em_k3": (20000, [(0.5, 200, 20), (0.3, 420, 35), (0.2, 700, 60)], 3, 1),
    "seeds_em_k6": (60000, [(0.3, 150, 30), (0.25, 260, 25), (0.2, 400, 50), (0.15, 620, 40), (0.1, 900, 90)], 6, 2),
    "seeds_em_k2_small": (500, [(0.6, 50, 10), (0.4, 90, 15)], 2, 3),
    "seeds_em_k10_overlap": (40000, [(0.4, 300, 80), (0.6, 500, 120)], 10, 4),
}


def em_case(name):
    from sklearn.cluster import kmeans_plusplus
    from sklearn.mixture import GaussianMixture

    n, comps, k, seed = CASES[name]
    x = mixture(n, comps, seed)
    # the reference hands sklearn a torch tensor, which sklearn converts to float64 (see oracle/np_seeds.py)
    import torch

    xt = torch.from_numpy(x).reshape(-1, 1)
    _, indices = kmeans_plusplus(x[:, None].astype(np.float64), k, random_state=np.random.RandomState(100 + seed))

    class Injected(GaussianMixture):
        def _initialize_parameters(self, X, random_state, xp=None):
            resp = np.zeros((X.shape[0], self.n_components), dtype=X.dtype)
            resp[indices, np.arange(self.n_components)] = 1
            self._initialize(X, resp)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gm = Injected(n_components=k, n_init=1, init_params="k-means++")
        labels = gm.fit_predict(xt)
    np.savez_compressed(OUT / f"{name}.npz", x=x, indices=indices.astype(np.int64), weights=gm.weights_.astype(np.float64), means=gm.means_[:, 0].astype(np.float64),
                        covariances=gm.covariances_[:, 0, 0].astype(np.float64), n_iter=np.int64(gm.n_iter_), converged=np.bool_(gm.converged_),
                        lower_bound=np.float64(gm.lower_bound_), labels=labels.astype(np.uint8))
    print(name, "n_iter", gm.n_iter_, "converged", gm.converged_, "means", np.sort(gm.means_[:, 0]).round(1))


def split_case():
    """Reference split_lables on a 40^3 synthetic subject, subclasses 1 and 3, global numpy RNG seeded."""
    import importlib.util

    import torch

    import ref_import
    from fetalsyngen_b200.utils.phantom import label_phantom

    ref_import.load_reference()
    argv = sys.argv
    sys.argv = ["generate_seeds.py", "--bids_path", "/nonexistent", "--out_path", "/nonexistent", "--annotation", "feta"]
    try:
        spec = importlib.util.spec_from_file_location("ref_generate_seeds", ref_import.REF_ROOT / "scripts" / "generate_seeds.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv

    shape = (40, 40, 40)
    seg, _ = label_phantom(shape, seed=5)
    rs = np.random.RandomState(11)
    base = np.array([0, 900, 300, 450, 820, 520, 330, 480], dtype=np.float32)[seg]
    tex = rs.randn(*shape).astype(np.float32)
    image = np.maximum(base + 40 * tex + 60 * np.sin(np.arange(shape[2], dtype=np.float32) / 3)[None, None, :], 0).astype(np.float32)
    # non-brain tissue around the labelled region (becomes meta-label 4), zeros further out, a few NaNs
    g = np.meshgrid(*[np.linspace(-1, 1, s) for s in shape], indexing="ij")
    r = np.sqrt(sum(gi**2 for gi in g))
    image[(seg == 0) & (r > 0.85)] = 0
    image[(seg == 0) & (r <= 0.85)] = (200 + 70 * tex)[(seg == 0) & (r <= 0.85)].clip(1)
    image[3, 4, 5] = np.nan
    feta2meta = {1: 1, 4: 1, 2: 2, 6: 2, 5: 3, 7: 3, 3: 3}
    out = {"image": image, "seg": seg.astype(np.float32)}
    for sub in (1, 3):
        img_t = torch.from_numpy(np.nan_to_num(image, nan=0.0))[None]  # process_subject: NaN -> 0, unsqueeze(0)
        seg_t = torch.from_numpy(seg.astype(np.float32))[None]
        np.random.seed(77)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = mod.split_lables(image=img_t, segmentation=seg_t, subclasses=sub, feta2meta=feta2meta)
        for m, vol in res[sub].items():
            out[f"sub{sub}_m{m}"] = np.asarray(vol)[0].astype(np.int8)
    np.savez_compressed(OUT / "seeds_split.npz", **out)
    print("seeds_split", {k: np.unique(v).tolist() for k, v in out.items() if k.startswith("sub")})


if __name__ == "__main__":
    for name in CASES:
        em_case(name)
    split_case()
