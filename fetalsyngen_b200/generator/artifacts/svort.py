"""Host side of the slice-acquisition simulator: rigid slice transforms, motion trajectories,
interleaving and the acquisition PSF.  Mirrors the reference's ``svort`` helpers
(``fetalsyngen/generator/artifacts/svort``: ``transform/transform.py:14-199,359-390``,
``transform/transform_convert.py:24-161`` (= the ``*_forward`` CUDA kernels of
``transform_convert_cuda_kernel.cu:14-65,190-264``), ``data/fetal_motion.py:22-48``,
``data/utils.py:18-102``).

These are <= 250 x 12 floats per sample: they stay on the host in float32 numpy and travel to
the GPU as one small upload per stack; nothing here is differentiated, so the reference's
autograd wrappers and backward kernels have no counterpart.  Random draws use the numpy global
RNG in the reference's order.
"""
from __future__ import annotations

from math import log, sqrt
from pathlib import Path

import numpy as np
from scipy.spatial.transform import Rotation

F32 = np.float32
TRANSFORM_EPS = 1e-6
GAUSSIAN_FWHM = 1 / (2 * sqrt(2 * log(2)))
SINC_FWHM = 1.206709128803223 * GAUSSIAN_FWHM


# ----------------------------------------------------------------------------- conversions
def axisangle2mat(ax: np.ndarray) -> np.ndarray:
    """(n,6) rotation vector + translation -> (n,3,4); Rodrigues with the small-angle branch."""
    ax = np.asarray(ax, dtype=F32)
    x, y, z = ax[:, 0], ax[:, 1], ax[:, 2]
    th2 = x * x + y * y + z * z
    big = th2 > TRANSFORM_EPS
    th = np.sqrt(np.where(big, th2, F32(1)))
    xn, yn, zn = x / th, y / th, z / th
    s, c = np.sin(th), np.cos(th)
    o = F32(1) - c
    mat = np.empty((ax.shape[0], 3, 4), dtype=F32)
    rows = ((c + xn * xn * o, xn * yn * o - zn * s, yn * s + xn * zn * o),
            (zn * s + xn * yn * o, c + yn * yn * o, -xn * s + yn * zn * o),
            (-yn * s + xn * zn * o, xn * s + yn * zn * o, c + zn * zn * o))
    one = np.ones_like(x)
    small = ((one, -z, y), (z, one, -x), (-y, x, one))
    for a in range(3):
        for b in range(3):
            mat[:, a, b] = np.where(big, rows[a][b], small[a][b])
    mat[:, :, 3] = ax[:, 3:]
    return mat


def mat2axisangle(mat: np.ndarray) -> np.ndarray:
    """(n,3,4) -> (n,6) through the 4-branch quaternion extraction."""
    m = np.asarray(mat, dtype=F32)
    r = [[m[:, a, b] for b in range(3)] for a in range(3)]
    d2, d01, d0n1 = r[2][2] < TRANSFORM_EPS, r[0][0] > r[1][1], r[0][0] < -r[1][1]
    one = F32(1)
    with np.errstate(invalid="ignore", divide="ignore"):
        s1 = F32(2) * np.sqrt(r[0][0] + r[1][1] + r[2][2] + one)
        s2 = F32(2) * np.sqrt(r[0][0] - r[1][1] - r[2][2] + one)
        s3 = F32(2) * np.sqrt(r[1][1] - r[0][0] - r[2][2] + one)
        s4 = F32(2) * np.sqrt(r[2][2] - r[0][0] - r[1][1] + one)
        q1 = (F32(0.25) * s1, (r[2][1] - r[1][2]) / s1, (r[0][2] - r[2][0]) / s1, (r[1][0] - r[0][1]) / s1)
        q2 = ((r[2][1] - r[1][2]) / s2, F32(0.25) * s2, (r[0][1] + r[1][0]) / s2, (r[0][2] + r[2][0]) / s2)
        q3 = ((r[0][2] - r[2][0]) / s3, (r[0][1] + r[1][0]) / s3, F32(0.25) * s3, (r[1][2] + r[2][1]) / s3)
        q4 = ((r[1][0] - r[0][1]) / s4, (r[0][2] + r[2][0]) / s4, (r[1][2] + r[2][1]) / s4, F32(0.25) * s4)
    c1, c2, c3 = (~d2) & (~d0n1), d2 & d01, d2 & (~d01)
    w, x, y, z = (np.where(c1, q1[k], np.where(c2, q2[k], np.where(c3, q3[k], q4[k]))).astype(F32) for k in range(4))
    neg = w < 0
    w, x, y, z = (np.where(neg, -v, v) for v in (w, x, y, z))
    tmp = x * x + y * y + z * z
    si = np.sqrt(tmp)
    theta = F32(2) * np.arctan2(si, w)
    with np.errstate(invalid="ignore", divide="ignore"):
        fac = np.where(tmp > TRANSFORM_EPS, theta / si, F32(2) / w).astype(F32)
    out = np.empty((m.shape[0], 6), dtype=F32)
    out[:, 0], out[:, 1], out[:, 2] = x * fac, y * fac, z * fac
    out[:, 3:] = m[:, :, 3]
    return out


def _bmm(a, b):
    return np.einsum("nij,njk->nik", a, b).astype(F32)


class RigidTransform:
    """n rigid transforms stored as axis-angle (n,6) or matrices (n,3,4) (transform.py:14-128)."""

    def __init__(self, data, trans_first=True):
        data = np.asarray(data, dtype=F32)
        self.trans_first = trans_first
        self._axisangle = self._matrix = None
        if data.shape[1] == 6:
            self._axisangle = data
        elif data.shape[1] == 3:
            self._matrix = data
        else:
            raise Exception("Unknown format for rigid transform!")

    def matrix(self, trans_first=True):
        mat = self._matrix if self._matrix is not None else axisangle2mat(self._axisangle)
        if self.trans_first and not trans_first:
            mat = np.concatenate([mat[:, :, :3], _bmm(mat[:, :, :3], mat[:, :, 3:])], -1)
        elif not self.trans_first and trans_first:
            mat = np.concatenate([mat[:, :, :3], _bmm(mat[:, :, :3].transpose(0, 2, 1), mat[:, :, 3:])], -1)
        return mat

    def axisangle(self, trans_first=True):
        if self.trans_first == trans_first:
            return self._axisangle.copy() if self._axisangle is not None else mat2axisangle(self._matrix)
        return mat2axisangle(self.matrix(trans_first))

    def compose(self, other):
        m1, m2 = self.matrix(True), other.matrix(True)
        R1, t1, R2, t2 = m1[:, :, :3], m1[:, :, 3:], m2[:, :, :3], m2[:, :, 3:]
        if R1.shape[0] != R2.shape[0]:
            n = max(R1.shape[0], R2.shape[0])
            R1, t1, R2, t2 = (np.broadcast_to(v, (n, *v.shape[1:])) for v in (R1, t1, R2, t2))
        return RigidTransform(np.concatenate([_bmm(R1, R2), t2 + _bmm(R2.transpose(0, 2, 1), t1)], -1), True)

    def __getitem__(self, idx):
        data = self._axisangle if self._axisangle is not None else self._matrix
        data = data[idx]
        if data.ndim < (2 if self._axisangle is not None else 3):
            data = data[None]
        return RigidTransform(data, self.trans_first)

    def __len__(self):
        return (self._axisangle if self._axisangle is not None else self._matrix).shape[0]

    @staticmethod
    def cat(transforms):
        return RigidTransform(np.concatenate([t.matrix(True) for t in transforms], 0), True)


def mat_update_resolution(mat, res_from, res_to):
    fac = np.ones((1, 1, 4), dtype=F32)
    fac[..., 3] = res_from / res_to
    return (mat * fac).astype(F32)


# ----------------------------------------------------------------------------- random transforms
def random_angle(n, restricted):
    """transform.py:178-188 (numpy draws: a, b, c)."""
    a = 2 * np.pi * np.random.rand(n)
    b = np.arccos(2 * np.random.rand(n) - 1)
    c = np.pi * np.random.rand(n) if restricted else np.pi * (2 * np.random.rand(n) - 1)
    return Rotation.from_euler("ZXZ", np.stack([a, b, c], -1)).as_rotvec().astype(F32)


def random_init_stack_transforms(n_slice, gap, restricted, txy):
    """transform.py:359-369."""
    angle = np.broadcast_to(random_angle(1, restricted), (n_slice, 3))
    tz = ((np.arange(0, n_slice, dtype=F32) - F32((n_slice - 1) / 2.0)) * F32(gap)).astype(F32)
    if txy:
        tx = np.ones_like(tz) * np.random.uniform(-txy, txy)
        ty = np.ones_like(tz) * np.random.uniform(-txy, txy)
    else:
        tx = ty = np.zeros_like(tz)
    return RigidTransform(np.concatenate([angle, np.stack([tx, ty, tz], -1).astype(F32)], -1), True)


def reset_transform(transform: RigidTransform) -> RigidTransform:
    """transform.py:386-390: keep only the (centred) through-plane translation."""
    ax = transform.axisangle()
    ax[:, :-1] = 0
    ax[:, -1] -= ax[:, -1].mean(dtype=F32)
    return RigidTransform(ax)


def interleave_index(N, n_i):
    idx, t = [None] * N, 0
    for i in range(n_i):
        for j in range(i, N, n_i):
            idx[j] = t
            t += 1
    return idx


# ----------------------------------------------------------------------------- motion trajectories
_TRAJ = None
TRAJ_PATH = Path(__file__).resolve().parent / "traj_knots.npz"


def get_trajectory():
    """Knots of the reference's 154 rotation + 154 translation trajectories (linear interpolants on
    integer knots; converted from svort/data/traj.npy by tools/convert_traj.py)."""
    global _TRAJ
    if _TRAJ is None:
        d = np.load(TRAJ_PATH)
        _TRAJ = {k: d[k] for k in d.files}
    return _TRAJ


def _traj_eval(y, t):
    """scipy ``interp1d(kind='linear', fill_value='extrapolate')`` on knots 0..len(y)-1."""
    t = np.asarray(t, dtype=np.float64)
    hi = np.clip(np.searchsorted(np.arange(len(y), dtype=np.float64), t), 1, len(y) - 1)
    lo = hi - 1
    slope = (y[hi] - y[lo]) / (hi - lo).astype(np.float64)[:, None]
    return slope * (t - lo)[:, None] + y[lo]


def sample_motion(ts, rand=True) -> RigidTransform:
    """fetal_motion.py:22-48."""
    tr = get_trajectory()
    out = []
    for name in ("rot", "trans"):
        k = np.random.choice(len(tr[f"{name}_T"]))
        y = tr[f"{name}_y"][tr[f"{name}_off"][k] : tr[f"{name}_off"][k + 1]]
        T, dT = tr[f"{name}_T"][k], tr[f"{name}_dT"][k]
        t0 = np.random.uniform(0, T - ts[-1] / dT) if rand else 0
        v = _traj_eval(y, t0 + ts / dT)
        if rand:
            v = v[:, np.random.permutation(3)]
            v = v * (2 * (np.random.rand(1, 3) < 0.5) - 1)
        out.append(v)
    R = Rotation.from_euler("xyz", out[0]).as_matrix().astype(F32)
    trans = out[1].astype(F32)
    R = np.matmul(R, R[0].T).astype(F32)
    trans = trans - trans[0]
    return RigidTransform(np.concatenate([R, trans[:, :, None]], -1), trans_first=False)


# ----------------------------------------------------------------------------- PSF
def get_PSF(r_max=None, res_ratio=(1, 1, 3), threshold=1e-4) -> np.ndarray:
    """Anisotropic Gaussian PSF (depth, height, width), thresholded at 1e-4, cropped to its support
    and normalised (data/utils.py:61-102)."""
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


def psf_taps(psf: np.ndarray):
    """Compact list of the non-zero PSF entries in the reference kernels' loop order
    (slice_acq_cuda_kernel.cu:62-66): [ntaps,4] float32 (ix_p, iy_p, iz_p, value) and the largest
    tap offset length (the cull radius of ``fsg_slice_acq_*``)."""
    dp, hp, wp = psf.shape
    iz, iy, ix = np.meshgrid(np.arange(-(dp // 2), (dp + 1) // 2), np.arange(-(hp // 2), (hp + 1) // 2), np.arange(-(wp // 2), (wp + 1) // 2), indexing="ij")
    nz = psf != 0
    taps = np.stack([ix[nz], iy[nz], iz[nz], psf[nz]], -1).astype(F32)
    radius = float(np.sqrt((taps[:, :3] ** 2).sum(1).max())) + 1e-3 if len(taps) else 0.0
    return np.ascontiguousarray(taps), radius
