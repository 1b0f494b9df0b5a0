"""TEST INFRASTRUCTURE — numpy restatement of seed generation.
Only ``tests/``, ``smoke()`` and the CPU legs of the benchmarks may import this file; the product
never does.

Follows /root/reference/scripts/generate_seeds.py:
  :133-146  process_subject   NaN -> 0 in image and label, dhcp label 4 -> 0
  :175-187  subsplit_label    GaussianMixture(n_components, n_init=5, init_params="k-means++")
                              .fit_predict on the image voxels of one meta-label
  :190-211  split_lables      label -> meta-label LUT, (label == 0 & image != 0) -> 4,
                              one class per meta-label when subclasses == 1

The clustering itself lives in a third-party dependency that is not under /root/reference:
scikit-learn, pinned 1.6.1 (requirements.txt:61, environment.yml:179); the image here carries
1.9.0, whose GaussianMixture / kmeans_plusplus arithmetic on this path is the same.  Its published
algorithm is restated below for one feature and ``covariance_type="full"``, in float64: the reference
hands sklearn a *torch* tensor (monai MetaTensor, generate_seeds.py:176-181), whose dtype sklearn's
``validate_data(dtype=[float64, float32])`` does not recognise, so the float32 image is converted to
float64 and the whole fit runs in double precision (a float32 numpy array would have stayed float32 —
measured here: 13 of 2792 labels differ between the two on the golden subject):
  sklearn/mixture/_base.py          fit_predict loop, _e_step, _estimate_log_prob_resp,
                                    _initialize_parameters (k-means++ branch)
  sklearn/mixture/_gaussian_mixture.py  _estimate_gaussian_parameters, _estimate_gaussian_covariances_full,
                                    _compute_precision_cholesky, _estimate_log_gaussian_prob, _initialize, _m_step
  sklearn/cluster/_kmeans.py        _kmeans_plusplus (greedy D^2 seeding, 2 + int(log k) local trials)
Pinned against sklearn itself run in the build container: ``tests/golden/make_golden_seeds.py``
drives the unmodified ``GaussianMixture`` (a) with injected k-means++ indices and (b) end to end
with a seeded ``RandomState`` and stores inputs + results in ``tests/golden/seeds_*.npz``.
"""
from __future__ import annotations

import math

import numpy as np

F64 = np.float64
EPS10 = 10 * np.finfo(np.float64).eps
LOG_2PI = math.log(2 * math.pi)

FETA2META = {1: 1, 4: 1, 2: 2, 6: 2, 5: 3, 7: 3, 3: 3}           # generate_seeds.py:74
DHCP2META = {1: 1, 5: 1, 2: 2, 7: 2, 9: 2, 3: 3, 6: 3, 8: 3}     # generate_seeds.py:84


# ----------------------------------------------------------------------------- k-means++ seeding
def kmeans_plusplus(x: np.ndarray, k: int, rs: np.random.RandomState) -> np.ndarray:
    """Indices of the k seeds (_kmeans.py:_kmeans_plusplus with unit sample weights).  Squared
    distances are formed directly as (x - c)^2 (sklearn: |x|^2 - 2xc + |c|^2 clipped at 0)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    n = x.size
    trials = 2 + int(np.log(k))
    w = np.ones(n, dtype=np.float64)
    idx = np.full(k, -1, dtype=int)
    idx[0] = rs.choice(n, p=w / w.sum())

    def dist(c):
        return (x - c) ** 2

    closest = dist(x[idx[0]])
    pot = closest @ w
    for c in range(1, k):
        rv = rs.uniform(size=trials) * pot
        cand = np.searchsorted(np.cumsum(w * closest), rv)
        np.clip(cand, None, n - 1, out=cand)
        d = np.stack([np.minimum(closest, dist(x[i])) for i in cand])
        pots = d @ w
        b = int(np.argmin(pots))
        pot, closest, idx[c] = pots[b], d[b], cand[b]
    return idx


# ----------------------------------------------------------------------------- EM
def init_from_indices(x, indices, reg_covar=1e-6):
    """_initialize with the one-hot responsibilities of the k-means++ branch: component j owns the
    single sample indices[j].  Returns (weights, means, covariances)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    n = x.size
    nk = 1.0 + EPS10
    xs = x[np.asarray(indices)]
    means = xs / nk
    diff = xs - means
    cov = (diff * diff) / nk + reg_covar
    weights = np.full(len(xs), nk / n, dtype=np.float64)
    return weights, means, cov


def weighted_log_prob(x, weights, means, cov):
    """_estimate_weighted_log_prob for one feature: [n, k]."""
    pc = 1.0 / np.sqrt(cov)  # precision Cholesky factor
    log_det = np.log(pc)
    y = x[:, None] * pc[None, :] - (means * pc)[None, :]
    lp = -0.5 * (LOG_2PI + y * y) + log_det[None, :]
    return lp + np.log(weights)[None, :]


def logsumexp_rows(a):
    m = a.max(axis=1, keepdims=True)
    return np.log(np.exp(a - m).sum(axis=1)) + m[:, 0]


def e_step(x, weights, means, cov):
    wlp = weighted_log_prob(x, weights, means, cov)
    lpn = logsumexp_rows(wlp)
    return float(lpn.mean()), wlp - lpn[:, None]


def m_step(x, log_resp, reg_covar=1e-6):
    resp = np.exp(log_resp)
    nk = resp.sum(axis=0) + EPS10
    means = (resp.T @ x) / nk
    cov = np.empty_like(means)
    for j in range(means.size):
        diff = x - means[j]
        cov[j] = ((resp[:, j] * diff) @ diff) / nk[j] + reg_covar
    return nk / nk.sum(), means, cov


def fit_from_indices(x, indices, max_iter=100, tol=1e-3, reg_covar=1e-6):
    """One initialisation of BaseMixture.fit_predict: EM from the given seeds until
    |delta lower bound| < tol.  Returns a dict with the parameters after the last M step."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    w, mu, cov = init_from_indices(x, indices, reg_covar)
    lower, converged, trace, n_iter = -np.inf, False, [], 0
    for n_iter in range(1, max_iter + 1):
        prev = lower
        lower, log_resp = e_step(x, w, mu, cov)
        w, mu, cov = m_step(x, log_resp, reg_covar)
        trace.append(float(lower))
        if abs(lower - prev) < tol:
            converged = True
            break
    return {"weights": w, "means": mu, "covariances": cov, "n_iter": n_iter, "converged": converged, "lower_bound": float(lower), "trace": np.asarray(trace)}


def predict(x, fit):
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    return np.argmax(weighted_log_prob(x, fit["weights"], fit["means"], fit["covariances"]), axis=1)


def fit_predict(x, k, rs: np.random.RandomState, n_init=5, max_iter=100, tol=1e-3, reg_covar=1e-6, return_fit=False):
    """GaussianMixture(k, n_init=5, init_params="k-means++", random_state=rs).fit_predict(x[:, None])."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    if x.size < k:
        raise ValueError(f"Expected n_samples >= n_components but got n_components = {k}, n_samples = {x.size}")
    best = None
    for _ in range(n_init):
        f = fit_from_indices(x, kmeans_plusplus(x, k, rs), max_iter, tol, reg_covar)
        if best is None or f["lower_bound"] > best["lower_bound"]:
            best = f
    labels = predict(x, best)
    return (labels, best) if return_fit else labels


# ----------------------------------------------------------------------------- label handling
def meta_labels(image, segmentation, annotation="feta"):
    """generate_seeds.py:133-146,190-198: uint8 meta-label volume (0 background, 1 CSF, 2 GM, 3 WM, 4 non-brain)."""
    image = np.nan_to_num(np.asarray(image, dtype=np.float32), nan=0.0, posinf=np.inf, neginf=-np.inf)
    seg = np.nan_to_num(np.asarray(segmentation, dtype=np.float32), nan=0.0, posinf=np.inf, neginf=-np.inf)
    table = FETA2META
    if annotation == "dhcp":
        seg = np.where(seg == 4, 0, seg)
        table = DHCP2META
    meta = np.zeros(seg.shape, dtype=np.uint8)
    for lab, m in table.items():
        meta[seg == lab] = m
    meta[(seg == 0) & (image != 0)] = 4
    return image, meta


def split_labels(image, segmentation, subclasses, rs, annotation="feta", n_init=5):
    """generate_seeds.py:190-211: {meta-label: int8 seed volume} for one number of subclasses."""
    image, meta = meta_labels(image, segmentation, annotation)
    out = {}
    for m in range(1, 5):
        mask = meta == m
        vol = np.zeros(meta.shape, dtype=np.int8)
        if subclasses == 1:
            vol[mask] = 10 * m
        else:
            vol[mask] = fit_predict(image[mask], subclasses, rs, n_init=n_init) + 10 * m
        out[m] = vol
    return out
