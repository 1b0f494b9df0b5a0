"""TEST INFRASTRUCTURE — numpy restatement of the native kernels behind the SimulateMotion
artifact and of the small host formulas around them.  Only ``tests/``, ``smoke()`` and the CPU
legs of ``bench.py`` may import this file; the product never does.

Pinned (a) on the GPU box against the reference's own extension built from the reference's
sources into ``oracle/_ref`` (``tests/test_gpu_motion.py``) and (b) in the build container as the
kernel stand-in under the *unmodified* reference ``Scanner`` / ``PSFReconstructor`` when the
golden vectors ``tests/golden/motion_*.npz`` are generated (``tests/golden/make_golden_motion.py``).

Follows (paths relative to /root/reference/fetalsyngen/generator/artifacts/svort):
  slice_acquisition/slice_acq_cuda_kernel.cu:17-171   forward, interp_psf = false, no masks
  slice_acquisition/slice_acq_cuda_kernel.cu:472-693  adjoint, interp_psf = true, + equalize
  transform/transform_convert_cuda_kernel.cu:14-65    axisangle2mat
  transform/transform_convert_cuda_kernel.cu:190-264  mat2axisangle
  data/utils.py:18-27,61-102                          interleave_index, get_PSF
All arithmetic is float32 with the kernels' double intermediates where the C source has a
``/ 2.`` literal.  (The reference is compiled with FMA contraction, numpy rounds every product:
agreement is to float tolerance, not bit-exact.)
"""
from __future__ import annotations

from math import log, sqrt

import numpy as np

F32 = np.float32
GAUSSIAN_FWHM = 1 / (2 * sqrt(2 * log(2)))
SINC_FWHM = 1.206709128803223 * GAUSSIAN_FWHM
TRANSFORM_EPS = 1e-6


# ----------------------------------------------------------------------------- small host formulas
def interleave_index(N, n_i):
    """data/utils.py:18-27."""
    idx, t = [None] * N, 0
    for i in range(n_i):
        for j in range(i, N, n_i):
            idx[j] = t
            t += 1
    return idx


def get_psf(r_max=None, res_ratio=(1, 1, 3), threshold=1e-4):
    """Anisotropic Gaussian PSF, thresholded, cropped, normalised (data/utils.py:61-102)."""
    sx, sy, sz = SINC_FWHM * res_ratio[0], SINC_FWHM * res_ratio[1], GAUSSIAN_FWHM * res_ratio[2]
    if r_max is None:
        r_max = max(max(int(2 * r + 1) for r in (sx, sy, sz)), 4)
    x = np.linspace(-r_max, r_max, 2 * r_max + 1, dtype=F32)
    gz, gy, gx = np.meshgrid(x, x, x, indexing="ij")
    psf = np.exp(F32(-0.5) * (gx**2 / F32(sx**2) + gy**2 / F32(sy**2) + gz**2 / F32(sz**2))).astype(F32)
    psf[np.abs(psf) < threshold] = 0
    rx = int(np.nonzero(psf.sum((0, 1)) > 0)[0][0])
    ry = int(np.nonzero(psf.sum((0, 2)) > 0)[0][0])
    rz = int(np.nonzero(psf.sum((1, 2)) > 0)[0][0])
    n = 2 * r_max + 1
    psf = np.ascontiguousarray(psf[rz : n - rz, ry : n - ry, rx : n - rx])
    return (psf / psf.sum()).astype(F32)


def axisangle2mat(ax):
    """(n,6) axis-angle + translation -> (n,3,4) (transform_convert_cuda_kernel.cu:14-65)."""
    ax = np.asarray(ax, dtype=F32)
    n = ax.shape[0]
    mat = np.zeros((n, 3, 4), dtype=F32)
    for i in range(n):
        x, y, z = ax[i, :3]
        th2 = x * x + y * y + z * z
        if th2 > TRANSFORM_EPS:
            th = np.sqrt(th2)
            x, y, z = x / th, y / th, z / th
            s, c = np.sin(th), np.cos(th)
            o = F32(1) - c
            mat[i, :, :3] = [[c + x * x * o, x * y * o - z * s, y * s + x * z * o],
                             [z * s + x * y * o, c + y * y * o, -x * s + y * z * o],
                             [-y * s + x * z * o, x * s + y * z * o, c + z * z * o]]
        else:
            mat[i, :, :3] = [[1, -z, y], [z, 1, -x], [-y, x, 1]]
        mat[i, :, 3] = ax[i, 3:]
    return mat


def mat2axisangle(mat):
    """(n,3,4) -> (n,6) through a quaternion (transform_convert_cuda_kernel.cu:190-264)."""
    mat = np.asarray(mat, dtype=F32)
    out = np.zeros((mat.shape[0], 6), dtype=F32)
    for i, m in enumerate(mat):
        r00, r01, r02, r10, r11, r12, r20, r21, r22 = (m[a, b] for a in range(3) for b in range(3))
        d2, d01, d0n1 = r22 < TRANSFORM_EPS, r00 > r11, r00 < -r11
        if not d2 and not d0n1:
            s = F32(2) * np.sqrt(r00 + r11 + r22 + F32(1))
            w, x, y, z = F32(0.25) * s, (r21 - r12) / s, (r02 - r20) / s, (r10 - r01) / s
        elif d2 and d01:
            s = F32(2) * np.sqrt(r00 - r11 - r22 + F32(1))
            w, x, y, z = (r21 - r12) / s, F32(0.25) * s, (r01 + r10) / s, (r02 + r20) / s
        elif d2 and not d01:
            s = F32(2) * np.sqrt(r11 - r00 - r22 + F32(1))
            w, x, y, z = (r02 - r20) / s, (r01 + r10) / s, F32(0.25) * s, (r12 + r21) / s
        else:
            s = F32(2) * np.sqrt(r22 - r00 - r11 + F32(1))
            w, x, y, z = (r10 - r01) / s, (r02 + r20) / s, (r12 + r21) / s, F32(0.25) * s
        if w < 0:
            w, x, y, z = -w, -x, -y, -z
        tmp = x * x + y * y + z * z
        si = np.sqrt(tmp)
        theta = F32(2) * np.arctan2(si, w)
        fac = theta / si if tmp > TRANSFORM_EPS else F32(2.0) / w
        out[i, :3] = [x * fac, y * fac, z * fac]
        out[i, 3:] = m[:, 3]
    return out


# ----------------------------------------------------------------------------- kernel geometry
def _centers(T, h, w, D, H, W, res):
    """World centre of every slice pixel: (n,h,w) float32 x/y/z (kernel lines 41-57)."""
    T = np.asarray(T, dtype=F32)
    res = np.float64(F32(res))
    ix = np.arange(w, dtype=np.float64)[None, None, :]
    iy = np.arange(h, dtype=np.float64)[None, :, None]
    t = T.astype(np.float64)
    _x = ((ix - (w - 1) / 2.0) * res + t[:, 0, 3][:, None, None]).astype(F32) + np.zeros((1, h, 1), F32)
    _y = ((iy - (h - 1) / 2.0) * res + t[:, 1, 3][:, None, None]).astype(F32) + np.zeros((1, 1, w), F32)
    _z = T[:, 2, 3][:, None, None] + np.zeros((1, h, w), F32)
    R = T[:, :, :3][:, :, :, None, None]
    xc = R[:, 0, 0] * _x + R[:, 0, 1] * _y + R[:, 0, 2] * _z
    yc = R[:, 1, 0] * _x + R[:, 1, 1] * _y + R[:, 1, 2] * _z
    zc = R[:, 2, 0] * _x + R[:, 2, 1] * _y + R[:, 2, 2] * _z
    xc = (xc.astype(np.float64) + (W - 1) / 2.0).astype(F32)
    yc = (yc.astype(np.float64) + (H - 1) / 2.0).astype(F32)
    zc = (zc.astype(np.float64) + (D - 1) / 2.0).astype(F32)
    return xc, yc, zc


def _taps(psf):
    """Non-zero PSF entries in the kernels' loop order: (ix_p, iy_p, iz_p, value)."""
    dp, hp, wp = psf.shape
    out = []
    for a, izp in enumerate(range(-(dp // 2), (dp + 1) // 2)):
        for b, iyp in enumerate(range(-(hp // 2), (hp + 1) // 2)):
            for c, ixp in enumerate(range(-(wp // 2), (wp + 1) // 2)):
                if psf[a, b, c] != 0:
                    out.append((ixp, iyp, izp, F32(psf[a, b, c])))
    return out


def psf_taps(psf) -> np.ndarray:
    """[ntaps,4] float32 table (ix_p, iy_p, iz_p, value) — the layout ``fsg_slice_acq_*`` take."""
    return np.asarray(_taps(np.asarray(psf, dtype=F32)), dtype=F32).reshape(-1, 4)


def _tap_pos(T, xc, yc, zc, ixp, iyp, izp):
    R = np.asarray(T, dtype=F32)[:, :, :3][:, :, :, None, None]
    fx, fy, fz = F32(ixp), F32(iyp), F32(izp)
    x = xc + R[:, 0, 0] * fx + R[:, 0, 1] * fy + R[:, 0, 2] * fz
    y = yc + R[:, 1, 0] * fx + R[:, 1, 1] * fy + R[:, 1, 2] * fz
    z = zc + R[:, 2, 0] * fx + R[:, 2, 1] * fy + R[:, 2, 2] * fz
    return x, y, z


def slice_acq_forward(transforms, vol, psf, slice_shape, res_slice):
    """slices (n,h,w): PSF-weighted trilinear samples of ``vol`` (D,H,W) — kernel lines 17-171,
    linear branch, vol_mask = slices_mask = NULL."""
    vol = np.asarray(vol, dtype=F32)
    psf = np.asarray(psf, dtype=F32)
    D, H, W = vol.shape
    h, w = slice_shape
    T = np.asarray(transforms, dtype=F32)
    xc, yc, zc = _centers(T, h, w, D, H, W, res_slice)
    val = np.zeros(xc.shape, F32)
    wsum = np.zeros(xc.shape, F32)
    flat = vol.reshape(-1)
    Sy, Sz = W, H * W
    for ixp, iyp, izp, pv in _taps(psf):
        x, y, z = _tap_pos(T, xc, yc, zc, ixp, iyp, izp)
        ok = ~((x < 0) | (y < 0) | (z < 0) | (x >= W - 1) | (y >= H - 1) | (z >= D - 1))
        if not ok.any():
            continue
        x, y, z = x[ok], y[ok], z[ok]
        xf, yf, zf = np.floor(x), np.floor(y), np.floor(z)
        wx, wy, wz = x - xf, y - yf, z - zf
        iv = zf.astype(np.int64) * Sz + yf.astype(np.int64) * Sy + xf.astype(np.int64)
        one = F32(1)
        v_acc, w_acc = val[ok], wsum[ok]
        for wgt, off in (((one - wx) * (one - wy) * (one - wz), 0), (wx * (one - wy) * (one - wz), 1), ((one - wx) * wy * (one - wz), Sy),
                         ((one - wx) * (one - wy) * wz, Sz), (wx * wy * (one - wz), 1 + Sy), (wx * (one - wy) * wz, 1 + Sz),
                         ((one - wx) * wy * wz, Sy + Sz), (wx * wy * wz, Sy + Sz + 1)):
            p = wgt * pv
            v_acc = v_acc + p * flat[iv + off]
            w_acc = w_acc + p
        val[ok], wsum[ok] = v_acc, w_acc
    out = np.zeros(xc.shape, F32)
    nz = wsum > 0
    out[nz] = val[nz] / wsum[nz]
    return out


def _c_round(v):
    """C ``round``: half away from zero (unlike numpy's half-to-even)."""
    return np.where(v >= 0, np.floor(v + F32(0.5)), np.ceil(v - F32(0.5))).astype(F32)


def slice_acq_adjoint(transforms, psf, slices, vol_shape, res_slice, equalize=True):
    """PSF reconstruction: scatter of the slices into a volume (kernel lines 472-693, NN branch with
    interpolated PSF, no masks) followed by ``vol /= vol_weight`` where the weight is positive."""
    psf = np.asarray(psf, dtype=F32)
    slices = np.asarray(slices, dtype=F32)
    D, H, W = vol_shape
    n, h, w = slices.shape
    T = np.asarray(transforms, dtype=F32)
    xc3, yc3, zc3 = _centers(T, h, w, D, H, W, res_slice)
    Sy, Sz = W, H * W
    taps = _taps(psf)
    n_of = np.broadcast_to(np.arange(n)[:, None, None], xc3.shape)
    weight = np.zeros(xc3.shape, F32)
    cache = []
    for ixp, iyp, izp, _ in taps:
        x, y, z = _tap_pos(T, xc3, yc3, zc3, ixp, iyp, izp)
        ok = ~((x < 0) | (y < 0) | (z < 0) | (x >= W - 1) | (y >= H - 1) | (z >= D - 1))
        if not ok.any():
            cache.append(None)
            continue
        xr, yr, zr = _c_round(x[ok]), _c_round(y[ok]), _c_round(z[ok])
        # per-element slice index for the rotation lookup
        nn = n_of[ok]
        R = T[:, :, :3][nn]
        dx, dy, dz = xr - xc3[ok], yr - yc3[ok], zr - zc3[ok]
        dp, hp, wp = psf.shape
        xp = ((R[:, 0, 0] * dx + R[:, 1, 0] * dy + R[:, 2, 0] * dz).astype(np.float64) + (wp - 1) / 2.0).astype(F32)
        yp = ((R[:, 0, 1] * dx + R[:, 1, 1] * dy + R[:, 2, 1] * dz).astype(np.float64) + (hp - 1) / 2.0).astype(F32)
        zp = ((R[:, 0, 2] * dx + R[:, 1, 2] * dy + R[:, 2, 2] * dz).astype(np.float64) + (dp - 1) / 2.0).astype(F32)
        good = ~((xp < 0) | (yp < 0) | (zp < 0) | (xp >= wp - 1) | (yp >= hp - 1) | (zp >= dp - 1))
        xp, yp, zp = xp[good], yp[good], zp[good]
        xf, yf, zf = np.floor(xp), np.floor(yp), np.floor(zp)
        wx, wy, wz = xp - xf, yp - yf, zp - zf
        iv = zf.astype(np.int64) * wp * hp + yf.astype(np.int64) * wp + xf.astype(np.int64)
        q = psf.reshape(-1)
        one = F32(1)
        v = np.zeros(xp.shape, F32)
        v = v + (one - wx) * (one - wy) * (one - wz) * q[iv]
        v = v + wx * (one - wy) * (one - wz) * q[iv + 1]
        v = v + (one - wx) * wy * (one - wz) * q[iv + wp]
        v = v + (one - wx) * (one - wy) * wz * q[iv + wp * hp]
        v = v + wx * wy * (one - wz) * q[iv + 1 + wp]
        v = v + wx * (one - wy) * wz * q[iv + 1 + wp * hp]
        v = v + (one - wx) * wy * wz * q[iv + wp + wp * hp]
        v = v + wx * wy * wz * q[iv + wp + wp * hp + 1]
        pix = np.flatnonzero(ok.reshape(-1))[good]
        vox = (zr[good].astype(np.int64) * Sz + yr[good].astype(np.int64) * Sy + xr[good].astype(np.int64))
        wflat = weight.reshape(-1)
        wflat[pix] = wflat[pix] + v  # one tap contributes at most once per pixel
        cache.append((pix, vox, v))
    vol = np.zeros(D * H * W, F32)
    vw = np.zeros(D * H * W, F32)
    wflat = weight.reshape(-1)
    sflat = slices.reshape(-1)
    for item in cache:
        if item is None:
            continue
        pix, vox, v = item
        keep = ~(wflat[pix] < 0.5)
        pix, vox, v = pix[keep], vox[keep], v[keep]
        pv = v / wflat[pix]
        np.add.at(vol, vox, pv * sflat[pix])
        np.add.at(vw, vox, pv)
    if equalize:
        nz = vw > 0
        vol[nz] = vol[nz] / vw[nz]
    return vol.reshape(D, H, W), vw.reshape(D, H, W)


# ----------------------------------------------------------------------------- slice artifacts / recon tail
def slice_gamma(slices, gamma):
    """Scanner.random_gamma (simulate_reco.py:225-236): float64 exponent, result divided by its max."""
    s = 300.0 * (np.asarray(slices, dtype=F32).astype(np.float64) / 300.0) ** np.float64(gamma)
    return s / s.max()


def slice_rician(slices, threshold, sigma, noise1, noise2):
    """Scanner.add_noise (:248-257) with the N(0,1) draws given as full-size arrays."""
    s = np.array(slices, copy=True)
    m = s > threshold
    s[m] = np.sqrt((s[m] + noise1[m] * sigma) ** 2 + (noise2[m] * sigma) ** 2)
    return s


def signal_void_mask(h, w, yc, xc, theta, a, A, sx):
    """Multiplicative mask of Scanner.signal_void (:273-297) for one slice."""
    y = np.linspace(-(h - 1) / 2, (h - 1) / 2, h, dtype=F32)[:, None] - F32(yc)
    x = np.linspace(-(w - 1) / 2, (w - 1) / 2, w, dtype=F32)[None, :] - F32(xc)
    c, s = np.cos(F32(theta)), np.sin(F32(theta))
    x, y = c * x - s * y, s * x + c * y
    sy = F32(a) ** 2 / F32(sx)
    gx, gy = F32(-0.5) / F32(sx) ** 2, F32(-0.5) / sy**2
    return (F32(1) - F32(A) * np.exp(gx * x**2 + gy * y**2)).astype(F32)


def smooth3(vol):
    """3^3 mean with zero padding (PSFReconstructor.smooth_volume, :584-595)."""
    v = np.pad(np.asarray(vol, dtype=F32), 1)
    out = np.zeros(vol.shape, F32)
    D, H, W = vol.shape
    for dz in range(3):
        for dy in range(3):
            for dx in range(3):
                out += v[dz : dz + D, dy : dy + H, dx : dx + W] * F32(1 / 27)
    return out


def merge(weight, rec, gt):
    """merge_volumes (:692-709)."""
    return (weight * rec + (F32(1) - weight) * gt).astype(F32)
