"""TEST INFRASTRUCTURE — CPU restatement (numpy, float32) of FetalSynthGen's per-sample path.

This file is the *oracle*: a from-scratch restatement of the arithmetic of the reference's
generation path, written so that every float32 operation is a separately rounded IEEE op in
the reference's order (numpy never contracts to FMA).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import it — never the product path (``fetalsyngen_b200/``), which must fail loudly when the
CUDA library is missing.

Parity pin: ``tests/test_oracle_golden.py`` checks every function below against golden vectors
produced by running the unmodified reference in the build container
(``tests/golden/make_golden.py``), including two cases at the full 256^3 size on the reference's bundled
subjects (``tests/golden/full_*.npz``).  The reference itself ships no golden vectors / KATs for this path
(SURVEY.md §4).

All ``file:line`` citations are relative to the reference tree (``fetalsyngen/...``).

Conventions: volumes are C-contiguous ``[x, y, z]`` (z fastest), float32 unless said otherwise.
The only torch use is ``torch.arange`` / ``torch.linspace`` for the 1-D position tables,
because their float32 rounding is implementation-defined (SURVEY.md §7 "hard parts") and the
reference builds those tables with exactly these calls.
"""
from __future__ import annotations

import numpy as np
import torch

f32 = np.float32


# =============================================================================== affine (a3)
def make_affine_matrix(rot, sh, s) -> np.ndarray:
    """A = SHx @ SHy @ SHz @ Rx @ Ry @ Rz with rows scaled by s, float64.
    Follows utils/generation.py:39-71."""
    cx, cy, cz = np.cos(rot[0]), np.cos(rot[1]), np.cos(rot[2])
    sx, sy, sz = np.sin(rot[0]), np.sin(rot[1]), np.sin(rot[2])
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=np.float64)
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=np.float64)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=np.float64)
    shx = np.array([[1, 0, 0], [sh[1], 1, 0], [sh[2], 0, 1]], dtype=np.float64)
    shy = np.array([[1, sh[0], 0], [0, 1, 0], [0, sh[2], 1]], dtype=np.float64)
    shz = np.array([[1, 0, sh[0]], [0, 1, sh[1]], [0, 0, 1]], dtype=np.float64)
    a = shx @ shy @ shz @ rx @ ry @ rz
    for r in range(3):
        a[r, :] = a[r, :] * s[r]
    return a


# =============================================================================== zoom (a4, a10, a13)
def zoom_axis_table(n_in: int, factor: float):
    """1-D sampling table of ``myzoom_torch`` for one axis (utils/generation.py:315-361).
    Returns (n_out, floor_idx int32, ceil_idx int32, w_floor f32, w_ceil f32)."""
    factor = float(factor)
    delta = (1.0 - factor) / (2.0 * factor)
    n_out = int(np.round(n_in * factor))
    v = torch.arange(delta, delta + n_out / factor, 1 / factor, dtype=torch.float)[:n_out].numpy().copy()
    v[v < 0] = 0
    v[v > (n_in - 1)] = n_in - 1
    fl = np.floor(v).astype(np.int32)
    ce = np.minimum(fl + 1, n_in - 1).astype(np.int32)
    wc = (v - fl.astype(f32)).astype(f32)
    wf = (f32(1) - wc).astype(f32)
    return n_out, fl, ce, wf, wc


def zoom_linear(x: np.ndarray, factors) -> np.ndarray:
    """Separable linear resize x -> y -> z, each ``w_f*X[f] + w_c*X[c]`` rounded to f32
    (utils/generation.py:363-386).  ``x`` is [X,Y,Z] or [X,Y,Z,C]."""
    squeeze = x.ndim == 3
    if squeeze:
        x = x[..., None]
    x = x.astype(f32, copy=False)
    _, fx, cx, wfx, wcx = zoom_axis_table(x.shape[0], factors[0])
    _, fy, cy, wfy, wcy = zoom_axis_table(x.shape[1], factors[1])
    _, fz, cz, wfz, wcz = zoom_axis_table(x.shape[2], factors[2])
    t1 = wfx[:, None, None, None] * x[fx] + wcx[:, None, None, None] * x[cx]
    t2 = wfy[None, :, None, None] * t1[:, fy] + wcy[None, :, None, None] * t1[:, cy]
    y = wfz[None, None, :, None] * t2[:, :, fz] + wcz[None, None, :, None] * t2[:, :, cz]
    return y[..., 0] if squeeze else y


# =============================================================================== GMM (a2)
def tie_subclass_means(mus, seed_labels, generation_classes, perturb) -> np.ndarray:
    """mus[seed_labels] = clamp(mus[generation_classes] + 25*perturb, 0, 225)
    (generator/intensity/rand_gmm.py:139-145).  Right-hand side is evaluated before the write."""
    mus = mus.astype(f32).copy()
    if list(seed_labels) != list(generation_classes):
        rhs = mus[np.asarray(generation_classes)] + f32(25) * perturb.astype(f32)
        mus[np.asarray(seed_labels)] = np.clip(rhs, f32(0), f32(225))
    return mus


def gmm_intensities(labels: np.ndarray, mus, sigmas, noise) -> np.ndarray:
    """I = mus[L] + sigmas[L]*N ; I[I<0] = 0   (rand_gmm.py:146-149).  Label 0 is synthesised too."""
    lab = labels.astype(np.int64)
    out = mus.astype(f32)[lab] + sigmas.astype(f32)[lab] * noise.astype(f32)
    out[out < 0] = 0
    return out.astype(f32)


# =============================================================================== deformation (a5)
def deformation_coords(shape, size, A, c2, F):
    """Sample coordinates of the warp (generator/deformation/affine_nonrigid.py:64-84, 327-366).

    shape: image shape (mesh), size: generator shape (centre), A: 3x3, c2: 3, F: [X,Y,Z,3] or None.
    Returns (xx2, yy2, zz2) float32, clamped to [0, S-1] and shifted by floor(min)."""
    A = np.asarray(A, dtype=f32)
    c2 = np.asarray(c2, dtype=np.float64).astype(f32)  # 0-dim f64 operand is cast to f32 in-op
    c = ((np.asarray(size, dtype=np.float64) - 1) / 2).astype(f32)
    gx = np.arange(shape[0], dtype=f32)[:, None, None] - c[0]
    gy = np.arange(shape[1], dtype=f32)[None, :, None] - c[1]
    gz = np.arange(shape[2], dtype=f32)[None, None, :] - c[2]
    if F is not None:
        x1 = gx + F[..., 0]
        y1 = gy + F[..., 1]
        z1 = gz + F[..., 2]
    else:
        x1, y1, z1 = (np.broadcast_to(g, shape).astype(f32) for g in (gx, gy, gz))
    out = []
    for r in range(3):
        v = A[r, 0] * x1 + A[r, 1] * y1
        v = v + A[r, 2] * z1
        v = v + c2[r]
        v = v.astype(f32)
        v[v < 0] = 0
        v[v > (shape[r] - 1)] = shape[r] - 1
        lo = np.floor(v.min())
        v = v - f32(lo)
        out.append(v.astype(f32))
    return out[0], out[1], out[2]


# =============================================================================== interpolation (a6, a7)
def interp_nearest(x: np.ndarray, ii, jj, kk) -> np.ndarray:
    """round-half-even, clamp, gather (utils/generation.py:211-225)."""
    ir = np.clip(np.rint(ii).astype(np.int64), 0, x.shape[0] - 1)
    jr = np.clip(np.rint(jj).astype(np.int64), 0, x.shape[1] - 1)
    kr = np.clip(np.rint(kk).astype(np.int64), 0, x.shape[2] - 1)
    return x[ir, jr, kr]


def interp_linear(x: np.ndarray, ii, jj, kk) -> np.ndarray:
    """Trilinear, blend order x, y, z; 0 wherever a coordinate is <= 0 or > S-1
    (utils/generation.py:227-285)."""
    x = x.astype(f32, copy=False)
    sx, sy, sz = x.shape
    ok = (ii > 0) & (jj > 0) & (kk > 0) & (ii <= sx - 1) & (jj <= sy - 1) & (kk <= sz - 1)
    iv = np.where(ok, ii, f32(0)).astype(f32)
    jv = np.where(ok, jj, f32(0)).astype(f32)
    kv = np.where(ok, kk, f32(0)).astype(f32)
    fx = np.floor(iv).astype(np.int64)
    fy = np.floor(jv).astype(np.int64)
    fz = np.floor(kv).astype(np.int64)
    cx = np.minimum(fx + 1, sx - 1)
    cy = np.minimum(fy + 1, sy - 1)
    cz = np.minimum(fz + 1, sz - 1)
    wcx = iv - fx.astype(f32)
    wcy = jv - fy.astype(f32)
    wcz = kv - fz.astype(f32)
    wfx, wfy, wfz = f32(1) - wcx, f32(1) - wcy, f32(1) - wcz
    c00 = x[fx, fy, fz] * wfx + x[cx, fy, fz] * wcx
    c01 = x[fx, fy, cz] * wfx + x[cx, fy, cz] * wcx
    c10 = x[fx, cy, fz] * wfx + x[cx, cy, fz] * wcx
    c11 = x[fx, cy, cz] * wfx + x[cx, cy, cz] * wcx
    c0 = c00 * wfy + c10 * wcy
    c1 = c01 * wfy + c11 * wcy
    c = c0 * wfz + c1 * wcz
    return np.where(ok, c, f32(0)).astype(f32)


def apply_deformation(output, segmentation, coords, flip: bool, image=None):
    """Flip the *sources* along axis 0, then linear (image) / nearest (segmentation) sampling
    (affine_nonrigid.py:164-193).  ``coords`` None = deformation gate off."""
    if flip:
        output = output[::-1]
        segmentation = segmentation[::-1]
        image = image[::-1] if image is not None else None
    if coords is not None:
        ii, jj, kk = coords
        output = interp_linear(output, ii, jj, kk)
        segmentation = interp_nearest(segmentation, ii, jj, kk)
        if image is not None:
            image = interp_linear(image, ii, jj, kk)
    return np.ascontiguousarray(output), np.ascontiguousarray(segmentation), image


# =============================================================================== intensity augmentations
def gamma_transform(x, gamma) -> np.ndarray:
    """300 * (x/300)**gamma, float32 throughout (augmentation/synthseg.py:262-275)."""
    return (f32(300.0) * np.power((x.astype(f32) / f32(300.0)), f32(gamma))).astype(f32)


def bias_field(x, bf_low) -> np.ndarray:
    """x * exp(zoom(bf_low))  (synthseg.py:157-188); bf_low already multiplied by bf_std."""
    fac = np.asarray(x.shape, dtype=np.float64) / np.asarray(bf_low.shape, dtype=np.float64)
    return (x.astype(f32) * np.exp(zoom_linear(bf_low.astype(f32), fac))).astype(f32)


def gaussian_taps(sigma: float) -> np.ndarray:
    """Normalised taps exp(-(t/sigma)^2/2), t in [-ceil(3 sigma), ceil(3 sigma)]
    (utils/generation.py:74-81)."""
    sl = int(np.ceil(3 * sigma))
    ts = torch.linspace(-sl, sl, 2 * sl + 1, dtype=torch.float)
    g = torch.exp((-((ts / sigma) ** 2) / 2))
    return (g / g.sum()).numpy()


def blur_axis(x, taps, axis) -> np.ndarray:
    """Zero-padded 1-D correlation along ``axis`` (conv3d semantics, generation.py:88-109)."""
    r = len(taps) // 2
    pad = [(0, 0)] * 3
    pad[axis] = (r, r)
    xp = np.pad(x.astype(f32), pad)
    n = x.shape[axis]
    out = np.zeros_like(x, dtype=f32)
    for t, w in enumerate(taps):
        sl = [slice(None)] * 3
        sl[axis] = slice(t, t + n)
        out += f32(w) * xp[tuple(sl)]
    return out


def gaussian_blur_3d(x, stds) -> np.ndarray:
    out = x.astype(f32)
    for ax in range(3):
        if stds[ax] > 0:
            out = blur_axis(out, gaussian_taps(stds[ax]), ax)
    return out


def resample_axis_positions(n_in: int, res_in: float, spacing: float):
    """Low-res sample positions along one axis, float64 numpy -> float32
    (synthseg.py:84-102).  Returns (n_out, factor f64, positions f32)."""
    n_out = int(n_in * res_in / spacing)
    factor = n_out / n_in
    delta = (1.0 - factor) / (2.0 * factor)
    v = np.arange(delta, delta + n_out / factor, 1 / factor)[:n_out]
    return n_out, factor, v.astype(f32)


def resample_stds(spacing, res_in, blur_u) -> np.ndarray:
    """(0.85+0.3u)*ln5/pi*spacing/res, zeroed where spacing <= res (synthseg.py:78-80)."""
    spacing = np.asarray(spacing, dtype=np.float64)
    res_in = np.asarray(res_in, dtype=np.float64)
    stds = (0.85 + 0.3 * blur_u) * np.log(5) / np.pi * spacing / res_in
    stds[spacing <= res_in] = 0.0
    return stds


def downsample(x, res_in, spacing, stds):
    """Blur then trilinear resample onto the coarse grid (synthseg.py:63-107).
    Returns (low-res volume, factors f64[3])."""
    xb = gaussian_blur_3d(x, stds)
    tabs = [resample_axis_positions(x.shape[a], res_in[a], spacing[a]) for a in range(3)]
    ii, jj, kk = np.meshgrid(tabs[0][2], tabs[1][2], tabs[2][2], indexing="ij")
    lo = interp_linear(xb, ii.astype(f32), jj.astype(f32), kk.astype(f32))
    return lo, np.array([t[1] for t in tabs])


def add_noise(x, noise_std, noise) -> np.ndarray:
    """x + std*N, clamp >= 0 (synthseg.py:217-235)."""
    out = x.astype(f32) + f32(noise_std) * noise.astype(f32)
    out[out < 0] = 0
    return out.astype(f32)


def resize_back(x, factors) -> np.ndarray:
    """zoom by 1/factors then divide by the global max (synthseg.py:109-114)."""
    up = zoom_linear(x, 1 / np.asarray(factors, dtype=np.float64))
    return (up / up.max()).astype(f32)


def scale_intensity(x) -> np.ndarray:
    """monai ScaleIntensity(minv=0, maxv=1) as used at data/datasets.py:40,311."""
    lo, hi = x.min(), x.max()
    if lo == hi:
        return (x * f32(0)).astype(f32)
    return (((x - lo) / (hi - lo)) * f32(1.0) + f32(0.0)).astype(f32)


# =============================================================================== full base pipeline
def generate_base(labels, seg, p: dict, image=None):
    """Seed labels + segmentation + drawn parameters/noise -> (image f32, seg, stages dict).

    ``p`` holds everything the reference would have drawn (see tests/golden/make_golden.py for
    the key list); a missing / None key means that stage's gate was off.  Follows
    generator/model.py:94-229."""
    st = {}
    size = p.get("size", labels.shape)
    out = gmm_intensities(labels, p["mus"], p["sigmas"], p["gmm_noise"])
    st["intensity"] = out
    if p.get("A") is not None:
        F = None
        if p.get("Fsmall") is not None:
            fs = p["Fsmall"]
            F = zoom_linear(fs, np.asarray(labels.shape, dtype=np.float64) / np.asarray(fs.shape[:3], dtype=np.float64))
            st["F"] = F
        coords = deformation_coords(labels.shape, size, p["A"], p["c2"], F)
        st["coords"] = coords
    else:
        coords = None
    out, seg, image = apply_deformation(out, seg, coords, bool(p.get("flip", False)), image)
    st["warped"], st["seg"] = out, seg
    if p.get("gamma") is not None:
        out = gamma_transform(out, p["gamma"])
    st["gamma"] = out
    if p.get("bf_low") is not None:
        out = bias_field(out, p["bf_low"])
    st["bias"] = out
    factors = None
    if p.get("spacing") is not None:
        out, factors = downsample(out, p["resolution"], p["spacing"], p["stds"])
    st["lowres"] = out
    if p.get("noise_std") is not None:
        out = add_noise(out, p["noise_std"], p["noise"])
    st["noisy"] = out
    if factors is not None:
        out = resize_back(out, factors)
    st["final"] = out
    return out, seg, st


# =============================================================================== Philox4x32-10 (for the RNG KATs)
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr, key):
    """Counter-based Philox-4x32-10 (Salmon et al., SC'11).  ctr: 4 uint32, key: 2 uint32.
    The reference has no counter-based RNG; this restates the published algorithm so the CUDA
    generator can be checked bit-exactly against known-answer vectors."""
    c = [int(v) & 0xFFFFFFFF for v in ctr]
    k = [int(v) & 0xFFFFFFFF for v in key]
    for _ in range(10):
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + _W0) & 0xFFFFFFFF, (k[1] + _W1) & 0xFFFFFFFF]
    return c
