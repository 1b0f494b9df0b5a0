"""TEST INFRASTRUCTURE — CPU restatement (numpy, float32) of the SR-artifact stages of
FetalSynthGen (``fetalsyngen/generator/augmentation/artifacts.py`` and
``fetalsyngen/generator/artifacts/utils.py`` of the reference; ``file:line`` citations are
relative to the reference tree).  Same rules as ``np_oracle.py``: only tests, ``smoke()`` and the
CPU-baseline legs of ``bench.py`` may import this file.

Parity pin: ``tests/test_oracle_artifacts.py`` checks every function against golden vectors made
by running the unmodified reference classes (``tests/golden/make_golden_artifacts.py``).
Random draws are *inputs* here (the reference's captured tensors / parameters).
"""
from __future__ import annotations

import numpy as np

from np_oracle import gaussian_blur_3d

f32 = np.float32


# =============================================================================== MoG (a14, a17)
def mog_3d(shape, centers, sigmas) -> np.ndarray:
    """clamp(sum_k exp(-d_k^2 / 2), 0, 1) (generator/artifacts/utils.py:125-160).

    Quirk kept: a centre is unpacked as ``x0, y0, z0`` with **x = last axis** and z = first axis,
    i.e. ``center[0]`` moves the blob along axis 2.  ``sigmas``: scalar, per-centre scalar, or
    per-centre triple (sigma_x, sigma_y, sigma_z) in the same transposed convention."""
    D, H, W = shape
    z = np.arange(D, dtype=f32)[:, None, None]
    y = np.arange(H, dtype=f32)[None, :, None]
    x = np.arange(W, dtype=f32)[None, None, :]
    mog = np.zeros(shape, dtype=f32)
    if not isinstance(sigmas, (list, np.ndarray)):
        sigmas = [sigmas] * len(centers)
    for center, sigma in zip(centers, sigmas):
        if isinstance(sigma, (list, np.ndarray)) and np.ndim(sigma) > 0:
            sx, sy, sz = sigma[0], sigma[1], sigma[2]
        else:
            sx = sy = sz = sigma
        x0, y0, z0 = center
        # float32 tensors divided by python/numpy float64 scalars stay float32 in torch
        d = ((x - f32(x0)) / f32(sx)) ** 2 + ((y - f32(y0)) / f32(sy)) ** 2 + ((z - f32(z0)) / f32(sz)) ** 2
        mog += np.exp(-d / f32(2)).astype(f32)
    return np.clip(mog, 0, 1).astype(f32)


def cortex_prior(shape) -> np.ndarray:
    """Frontal-lobe prior of BlurCortex.blur_proba (augmentation/artifacts.py:63-81)."""
    x, y, z = shape
    return mog_3d(shape, [(0, y, z // 2), (x, y, z // 2)], [x // 5, y // 5])


def blur_cortex(image, centers, sigmas, std_blurs) -> np.ndarray:
    """x*(1-g) + blur(x)*g with g the MoG of the drawn blobs (artifacts.py:112-126)."""
    g = mog_3d(image.shape, centers, sigmas)
    xb = gaussian_blur_3d(image.astype(f32), std_blurs)
    return (image.astype(f32) * (f32(1) - g) + xb * g).astype(f32)


# =============================================================================== multi-scale noise (a15)
def _upsample_axis(x, n_out, axis):
    """1-D linear up-sampling with torch's align_corners=False convention."""
    n_in = x.shape[axis]
    scale = n_in / n_out
    src = np.maximum(scale * (np.arange(n_out, dtype=np.float64) + 0.5) - 0.5, 0).astype(f32)
    i0 = np.floor(src).astype(np.int64)
    i1 = np.minimum(i0 + 1, n_in - 1)
    l1 = (src - i0.astype(f32)).astype(f32)
    l0 = (f32(1) - l1).astype(f32)
    shp = [1, 1, 1]
    shp[axis] = n_out
    return (l0.reshape(shp) * np.take(x, i0, axis=axis) + l1.reshape(shp) * np.take(x, i1, axis=axis)).astype(f32)


def upsample_trilinear(x, shape) -> np.ndarray:
    """F.interpolate(mode='trilinear', align_corners=False) (artifacts.py:315-320)."""
    out = x.astype(f32)
    for ax in (2, 1, 0):
        out = _upsample_axis(out, shape[ax], ax)
    return out


def multiscale_noise(shape, randn_stages) -> np.ndarray:
    """Sum of Gaussian noise over nstages scales (artifacts.py:308-322), normalised by max|.|."""
    n = len(randn_stages)
    lr = np.zeros([s // 2 ** n for s in shape], dtype=f32)
    for k in range(n):
        nxt = [s // 2 ** (n - 1 - k) for s in shape]
        lr = (lr + randn_stages[k].astype(f32)).astype(f32)
        lr = upsample_trilinear(lr, nxt)
    return (lr / np.abs(lr).max()).astype(f32)


# =============================================================================== Perlin (a15, a16)
def _fade(t):
    return t * t * t * (t * (t * f32(6) - f32(15)) + f32(10))


def perlin_3d(shape, res, theta, phi) -> np.ndarray:
    """generate_perlin_noise_3d (generator/artifacts/utils.py:224-327) with the random gradient
    angles given (theta = 2 pi u1, phi = 2 pi u2 from the captured ``torch.rand`` draws u1, u2);
    tileable in all axes."""
    lin = [np.linspace(0, res[i], shape[i]).astype(f32) if False else _torch_linspace(res[i], shape[i]) for i in range(3)]
    g = np.stack(np.meshgrid(*lin, indexing="ij"), axis=-1).astype(f32)
    cell = np.floor(g).astype(np.int64)
    loc = (g - cell.astype(f32)).astype(f32)
    th = (f32(2) * f32(np.pi) * theta.astype(f32)).astype(f32)
    ph = (f32(2) * f32(np.pi) * phi.astype(f32)).astype(f32)
    grad = np.stack([np.sin(ph) * np.cos(th), np.sin(ph) * np.sin(th), np.cos(ph)], axis=-1).astype(f32)
    grad[-1, :, :] = grad[0, :, :]
    grad[:, -1, :] = grad[:, 0, :]
    grad[:, :, -1] = grad[:, :, 0]

    def get(ix, iy, iz):
        return grad[np.minimum(ix, res[0]), np.minimum(iy, res[1]), np.minimum(iz, res[2])]

    def dot(gv, ox, oy, oz):
        d = loc - np.array([ox, oy, oz], dtype=f32)
        return (gv * d).sum(-1).astype(f32)

    cx, cy, cz = cell[..., 0], cell[..., 1], cell[..., 2]
    n000 = dot(get(cx, cy, cz), 0, 0, 0)
    n100 = dot(get(cx + 1, cy, cz), 1, 0, 0)
    n010 = dot(get(cx, cy + 1, cz), 0, 1, 0)
    n110 = dot(get(cx + 1, cy + 1, cz), 1, 1, 0)
    n001 = dot(get(cx, cy, cz + 1), 0, 0, 1)
    n101 = dot(get(cx + 1, cy, cz + 1), 1, 0, 1)
    n011 = dot(get(cx, cy + 1, cz + 1), 0, 1, 1)
    n111 = dot(get(cx + 1, cy + 1, cz + 1), 1, 1, 1)
    t = _fade(loc)
    tx, ty, tz = t[..., 0], t[..., 1], t[..., 2]
    n00 = n000 * (1 - tx) + tx * n100
    n10 = n010 * (1 - tx) + tx * n110
    n01 = n001 * (1 - tx) + tx * n101
    n11 = n011 * (1 - tx) + tx * n111
    n0 = n00 * (1 - ty) + ty * n10
    n1 = n01 * (1 - ty) + ty * n11
    return (n0 * (1 - tz) + tz * n1).astype(f32)


def _torch_linspace(end, steps):
    import torch

    return torch.linspace(0, int(end), int(steps)).numpy()


def fractal_noise_3d(shape, res, thetas, phis, persistence=0.5, lacunarity=2, increase=0.0) -> np.ndarray:
    """generate_fractal_noise_3d (utils.py:330-388): octaves summed, then
    (n + increase - min) / (max - min) clamped to [0, 1] (min/max of n *before* the increase)."""
    noise = np.zeros(shape, dtype=f32)
    freq, amp = 1, f32(1)
    for th, ph in zip(thetas, phis):
        noise = (noise + amp * perlin_3d(shape, (freq * res[0], freq * res[1], freq * res[2]), th, ph)).astype(f32)
        freq *= lacunarity
        amp = f32(amp * persistence)
    out = (noise + f32(increase) - noise.min()) / (noise.max() - noise.min())
    return np.clip(out, 0, 1).astype(f32)


def struct_noise(image, seg, lr_noise, noise_std, weight) -> np.ndarray:
    """Blend of the image with its noisy version inside seg > 0 (artifacts.py:323-339)."""
    x = image.astype(f32)
    noisy = np.clip(x + f32(noise_std) * lr_noise, 0, x.max() * f32(2)).astype(f32)
    mask = (seg > 0).astype(f32)
    mw = (mask * weight).astype(f32)
    return ((f32(1) - mw) * x + mw * noisy).astype(f32)


# =============================================================================== morphology (a17)
def _box_count(mask, k) -> np.ndarray:
    """Zero-padded k^3 box sum (apply_kernel, utils.py:163-171)."""
    r = k // 2
    out = mask.astype(np.int32)
    for ax in range(3):
        pad = [(0, 0)] * 3
        pad[ax] = (r, r)
        p = np.pad(out, pad)
        acc = np.zeros_like(out)
        n = out.shape[ax]
        for t in range(k):
            sl = [slice(None)] * 3
            sl[ax] = slice(t, t + n)
            acc += p[tuple(sl)]
        out = acc
    return out


def dilate(mask, k=3) -> np.ndarray:
    return (_box_count(mask, k) > 0).astype(np.uint8)  # utils.py:195-210


def erode(mask, k=3) -> np.ndarray:
    return (_box_count(mask, k) == k ** 3).astype(np.uint8)  # utils.py:174-192 (zero padding erodes at the border)


def build_halo(mask, radius) -> np.ndarray:
    """Dilation by skimage's ball(radius) as a 'same' zero-padded conv (artifacts.py:484-499)."""
    r = int(radius)
    m = mask.astype(bool)
    X, Y, Z = m.shape
    out = np.zeros_like(m)
    p = np.pad(m, r)
    for dx in range(-r, r + 1):
        for dy in range(-r, r + 1):
            rem = r * r - dx * dx - dy * dy
            if rem < 0:
                continue
            dz = int(np.floor(np.sqrt(rem)))
            # union over the z-run [-dz, dz]: a 1-D dilation of the shifted plane set
            sl = p[r + dx : r + dx + X, r + dy : r + dy + Y, :]
            cs = np.concatenate([np.zeros((X, Y, 1), dtype=np.int32), np.cumsum(sl, axis=2, dtype=np.int32)], axis=2)
            lo = np.arange(Z) + r - dz
            hi = np.arange(Z) + r + dz + 1
            out |= (cs[:, :, hi] - cs[:, :, lo]) > 0
    return out.astype(np.uint8)


def fuzzy_iteration(mask, perm) -> np.ndarray:
    """One generate_fuzzy_boundaries round (artifacts.py:501-522): 7^3 dilation ring, zero the
    first 90 % of the permuted ring voxels, keep voxels whose 3^3 neighbourhood holds more than 3
    survivors, then a 5^3 closing."""
    diff = (dilate(mask, 7).astype(np.int32) - mask.astype(np.int32)).astype(np.int32)
    nz = np.nonzero(diff)
    n = len(nz[0])
    idx = perm[: int(n * 0.9)]
    diff[nz[0][idx], nz[1][idx], nz[2][idx]] = 0
    dsamp = _box_count(diff, 3) > 3
    return erode(dilate(np.clip(mask.astype(np.int32) + dsamp, 0, 1), 5), 5)


def l1_dilations(mask, n) -> list:
    """[mask, mask, halo1(mask), halo1^2(mask), ...] of length n (artifacts.py:582-585);
    build_halo(., 1) is the 6-neighbour dilation."""
    stack = [mask.astype(np.uint8)] * 2
    for _ in range(n - 2):
        stack.append(build_halo(stack[-1], 1))
    return stack


def boundaries_mask(mask_halo, mask_modif, mog, n_generate_fuzzy):
    """Final mask of SimulatedBoundaries (artifacts.py:563-602)."""
    surf = (mask_modif.astype(np.int32) - mask_halo.astype(np.int32)) > 0
    surf_proba = np.where(surf, mog, f32(0)).astype(f32)
    n_dilate = 6 * (n_generate_fuzzy - 1)
    stack = np.stack(l1_dilations(mask_halo, n_dilate), 0) * mask_modif[None]
    idx = np.clip(np.rint(surf_proba * f32(len(stack)) - f32(1)).astype(np.int64), 0, None)
    return np.take_along_axis(stack, idx[None], 0)[0].astype(np.uint8), idx
