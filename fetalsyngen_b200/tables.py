"""Host-side 1-D tables that parameterise the kernels.

The float32 rounding of the reference's sample positions is implementation-defined
(``torch.arange(..., dtype=float32)`` in ``myzoom_torch``, utils/generation.py:318-338; numpy
float64 ``arange`` cast to float32 in ``RandResample``, synthseg.py:84-102), so the positions
are produced here with the very same library calls and handed to the kernels as
``fsg_tab {int16 f, int16 c, float wc}`` arrays.  The kernels then only do the blends.
Tables are tiny (<= a few KB), cached per key on the device.
"""
from __future__ import annotations

import numpy as np
import torch

TAB_DTYPE = np.dtype([("f", "<i2"), ("c", "<i2"), ("wc", "<f4")])


def _pack(v: np.ndarray, n_in: int, mark_outside: bool) -> np.ndarray:
    v = v.astype(np.float32)
    fl = np.floor(v)
    tab = np.zeros(v.shape[0], dtype=TAB_DTYPE)
    tab["f"] = fl.astype(np.int16)
    tab["c"] = np.minimum(fl.astype(np.int32) + 1, n_in - 1).astype(np.int16)
    tab["wc"] = v - fl
    if mark_outside:  # linear sampling returns 0 unless 0 < v <= n_in-1 (generation.py:229-235)
        out = ~((v > 0) & (v <= n_in - 1))
        tab["f"][out] = -1
        tab["c"][out] = 0
        tab["wc"][out] = 0
    return tab


def zoom_size(n_in: int, factor: float) -> int:
    return int(np.round(n_in * float(factor)))


def zoom_table(n_in: int, factor: float) -> np.ndarray:
    """Sampling table of ``myzoom_torch`` along one axis (utils/generation.py:315-361)."""
    factor = float(factor)
    delta = (1.0 - factor) / (2.0 * factor)
    n_out = zoom_size(n_in, factor)
    v = torch.arange(delta, delta + n_out / factor, 1 / factor, dtype=torch.float)[:n_out].numpy().copy()
    v[v < 0] = 0
    v[v > (n_in - 1)] = n_in - 1
    return _pack(v, n_in, mark_outside=False)


def resample_size(n_in: int, res_in: float, spacing: float) -> int:
    return int(n_in * res_in / spacing)


def resample_table(n_in: int, res_in: float, spacing: float):
    """Coarse-grid positions of ``RandResample`` along one axis (synthseg.py:84-102).
    Returns (table, factor)."""
    n_out = resample_size(n_in, res_in, spacing)
    factor = n_out / n_in
    delta = (1.0 - factor) / (2.0 * factor)
    v = np.arange(delta, delta + n_out / factor, 1 / factor)[:n_out]
    return _pack(v, n_in, mark_outside=True), factor


def gaussian_taps(sigma: float) -> np.ndarray:
    """``make_gaussian_kernel`` (utils/generation.py:74-81), evaluated on the host."""
    sl = int(np.ceil(3 * sigma))
    ts = torch.linspace(-sl, sl, 2 * sl + 1, dtype=torch.float)
    g = torch.exp((-((ts / sigma) ** 2) / 2))
    return (g / g.sum()).numpy()


def gaussian_taps_np(sigma: float) -> np.ndarray:
    """Same taps as ``gaussian_taps`` from the library (``fsg_gaussian_taps``: float32 arithmetic, no torch dispatch
    on the hot host path; the native step builder uses the same routine); differs from torch's by an ulp of exp()
    and of the normalising sum at most."""
    from . import _lib

    out = np.empty(2 * int(np.ceil(3 * sigma)) + 1, dtype=np.float32)
    n = _lib.load().fsg_gaussian_taps(float(sigma), out.ctypes.data, out.size)
    if n != out.size:
        raise _lib.FsgError(f"fsg_gaussian_taps({sigma}) returned {n}")
    return out


def _gaussian_taps_numpy(sigma: float) -> np.ndarray:
    sl = int(np.ceil(3 * sigma))
    ts = np.arange(-sl, sl + 1, dtype=np.float32)
    g = np.exp(-((ts / np.float32(sigma)) ** 2) / np.float32(2))
    return (g / g.sum(dtype=np.float32)).astype(np.float32)


def resample_stds(spacing, res_in, blur_u: float) -> np.ndarray:
    """Blur widths of the resolution simulation (synthseg.py:78-80)."""
    spacing = np.asarray(spacing, dtype=np.float64)
    res_in = np.asarray(res_in, dtype=np.float64)
    stds = (0.85 + 0.3 * blur_u) * np.log(5) / np.pi * spacing / res_in
    stds[spacing <= res_in] = 0.0
    return stds


def make_affine_matrix(rot, sh, s) -> np.ndarray:
    """3x3 float64 affine of the spatial deformation: shear_x . shear_y . shear_z . Rx . Ry . Rz,
    rows scaled (utils/generation.py:39-71).  Same six float64 matrices and the same left-to-right
    numpy products as the reference (bit-identical), built by index assignment into one array."""
    c0, c1, c2 = (float(v) for v in np.cos(np.asarray(rot, dtype=np.float64)))
    s0, s1, s2 = (float(v) for v in np.sin(np.asarray(rot, dtype=np.float64)))
    h0, h1, h2 = float(sh[0]), float(sh[1]), float(sh[2])
    m = np.zeros((6, 3, 3), dtype=np.float64)
    m[:, 0, 0] = m[:, 1, 1] = m[:, 2, 2] = 1.0
    m[0, 1, 0], m[0, 2, 0] = h1, h2
    m[1, 0, 1], m[1, 2, 1] = h0, h2
    m[2, 0, 2], m[2, 1, 2] = h0, h1
    m[3, 1, 1], m[3, 1, 2], m[3, 2, 1], m[3, 2, 2] = c0, -s0, s0, c0
    m[4, 0, 0], m[4, 0, 2], m[4, 2, 0], m[4, 2, 2] = c1, s1, -s1, c1
    m[5, 0, 0], m[5, 0, 1], m[5, 1, 0], m[5, 1, 1] = c2, -s2, s2, c2
    a = m[0]
    for k in range(1, 6):
        a = a @ m[k]
    return a * np.asarray(s, dtype=np.float64)[:, None]


class DeviceTables:
    """Per-device cache of uploaded tables / tap arrays (built lazily, keyed by their integer
    extents so a continuous random spacing still hits the cache)."""

    MAX_ENTRIES = 65536  # tiny arrays (<= a few KB each); the continuous blur width keys the tap arrays

    def __init__(self, device):
        self.device = torch.device(device)
        self._cache: dict = {}
        self.hold = False  # set by the engine while a launch batch is open: queued calls hold raw pointers into the cache

    def _put(self, key, make) -> torch.Tensor:
        t = self._cache.get(key)
        if t is None:
            arr = make()
            t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).copy()).to(self.device)
            if len(self._cache) > self.MAX_ENTRIES and not self.hold:
                # drop the tap arrays only (keyed by a continuous sigma); the integer-keyed tables are bounded by the extents
                for k in [k for k in self._cache if k[0] == "taps"]:
                    del self._cache[k]
            self._cache[key] = t
        return t

    def zoom(self, n_in: int, factor: float, expect: int | None = None) -> torch.Tensor:
        """Table of ``myzoom_torch`` along one axis; ``expect`` (checked when the table is built)
        is the output extent the caller relies on."""
        factor = float(factor)

        def make():
            if expect is not None and zoom_size(n_in, factor) != expect:
                raise ValueError(f"zoom of extent {n_in} by {factor} does not give {expect}")
            return zoom_table(n_in, factor)

        return self._put(("zoom", n_in, factor, expect), make)

    def resample(self, n_in: int, res_in: float, spacing: float):
        n_out = resample_size(n_in, res_in, spacing)
        key = ("resample", n_in, n_out)  # the positions only depend on the factor n_out / n_in
        t = self._cache.get(key)
        if t is None:
            tab, factor = resample_table(n_in, res_in, spacing)
            t = self._put(key, lambda: tab)
            self._cache[key + ("factor",)] = factor
        return t, self._cache[key + ("factor",)]

    def prewarm_resample(self, n_in: int, res_in: float, min_spacing: float, max_spacing: float) -> int:
        """Build, in one upload, every table the resolution simulation can ask for along an axis of
        extent ``n_in``: the down-sampling positions for each reachable coarse extent n_out and the
        ``myzoom_torch`` table that brings n_out back to n_in.  The spacing is continuous but the
        tables only depend on (n_in, n_out), so after this call the hot loop never builds a table."""
        lo = max(1, resample_size(n_in, res_in, max_spacing))
        hi = min(n_in, resample_size(n_in, res_in, min_spacing))
        todo = []
        for n_out in range(lo, hi + 1):
            kr = ("resample", n_in, n_out)
            if kr not in self._cache:
                factor = n_out / n_in
                delta = (1.0 - factor) / (2.0 * factor)
                v = np.arange(delta, delta + n_out / factor, 1 / factor)[:n_out]
                todo.append((kr, _pack(v, n_in, mark_outside=True), factor))
            back = float(1 / np.float64(n_out / n_in))
            kz = ("zoom", n_out, back, n_in)
            if kz not in self._cache and zoom_size(n_out, back) == n_in:
                todo.append((kz, zoom_table(n_out, back), None))
        if not todo:
            return 0
        blobs = [np.ascontiguousarray(t).view(np.uint8).reshape(-1) for _, t, _ in todo]
        offs = np.cumsum([0] + [(b.size + 15) // 16 * 16 for b in blobs])
        stage = np.zeros(int(offs[-1]), dtype=np.uint8)
        for b, o in zip(blobs, offs):
            stage[o : o + b.size] = b
        dev = torch.from_numpy(stage).to(self.device)
        for (key, _, factor), b, o in zip(todo, blobs, offs):
            self._cache[key] = dev[int(o) : int(o) + b.size]
            if factor is not None:
                self._cache[key + ("factor",)] = factor
        return len(todo)

    def taps(self, sigma: float) -> torch.Tensor:
        return self._put(("taps", float(sigma)), lambda: gaussian_taps(sigma))
