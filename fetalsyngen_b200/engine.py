"""Device engine: turns drawn per-sample parameters (``SamplePlan``) into batched launches of
the libfsg kernels.  torch is used for device memory and streams only.

Stage order and semantics follow ``FetalSynthGen.generate/augment``
(reference ``fetalsyngen/generator/model.py:94-229``):
  GMM -> warp(+flip, +gamma, +bias) -> blur -> down-sample(+noise) -> up-sample(/max)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from .tables import DeviceTables, gaussian_taps_np, resample_size, zoom_size

STAGE_GMM, STAGE_NOISE, STAGE_FIELD, STAGE_BIAS = 1, 2, 3, 4
# FSG_GMM_PAIRS=1: hand the GMM image to the warp fast path as 16-bit fixed-point z-pairs (one 32-bit gather
# per (x, y) row instead of two).  Measured r01h: warp 0.94 -> 0.79 ms, GMM 0.22 -> 0.29 ms per 8 volumes (net
# -4 % of the step) for a float error of ~1.5e-5 of the range instead of 3e-7: not worth it by default.
# FSG_GMM_PAIRS=2: float2 z-pairs (I[z], I[z+1]) — lossless (results identical to the plain path), one 8-byte
# gather per row in the warp; the GMM kernel writes 8 instead of 4 bytes per voxel.
_PAIRS_MODE = int(__import__("os").environ.get("FSG_GMM_PAIRS", "0") or 0)
_PAIRS = _PAIRS_MODE in (1, 2)
# FSG_WARP_TEX (default 1): samples that take the warp fast path get their GMM image written into a block-linear
# layered array (fsg_texvol) and gathered with two tld4 texture instructions per voxel instead of eight global
# loads; results are bit-identical to the linear path (FSG_WARP_TEX=0).
# The opt-in experiments on the linear source (FSG_GMM_PAIRS, FSG_WARP_TILE, FSG_WARP_PIPE) switch it off.
_env = __import__("os").environ
_TEX = (_env.get("FSG_WARP_TEX", "1") or "1") != "0" and not _PAIRS and _env.get("FSG_WARP_TILE", "0") in ("", "0") and _env.get("FSG_WARP_PIPE", "0") in ("", "0")


class TexVolume:
    """One block-linear intensity volume (``fsg_texvol``): CUDA array + texture + surface handles."""

    def __init__(self, shape):
        self.h = _lib.TexVol()
        _lib.call("fsg_texvol_create", int(shape[0]), int(shape[1]), int(shape[2]), C.byref(self.h))

    def upload(self, linear: torch.Tensor):
        _lib.call("fsg_texvol_copy", C.byref(self.h), linear.data_ptr(), 0, _stream())

    def download(self, linear: torch.Tensor):
        _lib.call("fsg_texvol_copy", C.byref(self.h), linear.data_ptr(), 1, _stream())

    def __del__(self):
        try:
            if self.h.array:
                _lib.call("fsg_texvol_destroy", C.byref(self.h))
        except Exception:
            pass


@dataclass
class SamplePlan:
    """Everything the reference would draw for one sample (None = that gate is off)."""

    mus: np.ndarray = None                # float32 [nlabels]
    sigmas: np.ndarray = None             # float32 [nlabels]
    gmm_noise: torch.Tensor | None = None  # injected N(0,1) draws [S], device (parity mode)
    rng_seed: int = 0                     # Philox key (production mode)
    sample_id: int = 0                    # Philox subsequence
    deform: bool = False
    flip: bool = False
    A: np.ndarray | None = None           # float32 3x3
    c2: np.ndarray | None = None          # float64 [3]
    center: np.ndarray | None = None      # float32 [3] = (size-1)/2
    fsmall: np.ndarray | None = None      # float32 [s0,s1,s2,3] (already scaled by nonlin_std)
    fsmall_dev: tuple | None = None       # ((s0,s1,s2), nonlin_std): draw the control grid on the device instead
    bf_dev: tuple | None = None           # ((b0,b1,b2), bf_std): same for the bias control grid
    gamma: float | None = None
    bf_low: np.ndarray | None = None      # float32 [b0,b1,b2] (already scaled by bf_std)
    spacing: np.ndarray | None = None     # float64 [3]
    stds: np.ndarray | None = None        # float64 [3]
    noise_std: float | None = None
    noise: torch.Tensor | None = None     # injected draws [n0*n1*n2] or [S] (parity mode)
    meta: dict = field(default_factory=dict)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(t: torch.Tensor, dtype, device, name):
    if t.device != device:
        raise ValueError(f"{name}: expected a tensor on {device}, got {t.device}")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return t


class SynthEngine:
    """Owns the per-device tables and scratch volumes for volumes of one shape."""

    def __init__(self, shape, resolution, device):
        self.shape = tuple(int(s) for s in shape)
        self.resolution = np.asarray(resolution, dtype=np.float64)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.FsgError(f"fetalsyngen_b200 runs on CUDA devices only (got device={device!r}); there is no CPU fallback")
        if not torch.cuda.is_available():
            raise _lib.FsgError("no CUDA device is available; fetalsyngen_b200 has no CPU fallback")
        _lib.load()
        self.nvox = int(np.prod(self.shape))
        self.tables = DeviceTables(self.device)
        self._scratch: dict = {}
        self._batch = None
        self.use_tex = _TEX
        self._texvols: list = []
        self._ptrs: dict = {}

    # ------------------------------------------------------------------ memory
    def scratch(self, name: str, batch: int, dtype=torch.float32, numel=None) -> torch.Tensor:
        """[batch, numel] scratch rows (grown on demand, never shrunk).  The row pitch is a multiple of 256 bytes so
        that every row is aligned for the vector kernels whatever the volume extents."""
        numel = self.nvox if numel is None else numel
        key = (name, dtype)
        t = self._scratch.get(key)
        if t is None or t.shape[0] < batch or t.shape[1] < numel:
            pitch = (numel + 63) // 64 * 64
            t = torch.empty((batch, pitch), dtype=dtype, device=self.device)
            self._scratch[key] = t
        return t[:, :numel] if t.shape[1] != numel else t

    def texvol(self, b: int) -> TexVolume:
        """The b-th block-linear intensity volume of this engine (created on first use)."""
        while len(self._texvols) <= b:
            self._texvols.append(TexVolume(self.shape))
        return self._texvols[b]

    # ---- raw device addresses of the cached tables (batch_step.py): one dictionary lookup per use
    def zoom_table_ptr(self, n_in: int, axis: int) -> int:
        """Table that zooms a control grid of extent n_in to the volume extent along `axis`."""
        key = ("z", n_in, axis)
        p = self._ptrs.get(key)
        if p is None:
            p = self._ptrs[key] = self.tables.zoom(n_in, self.shape[axis] / n_in, self.shape[axis]).data_ptr()
        return p

    def resample_table_ptr(self, axis: int, spacing: float):
        """(positions table address, factor) of the down-sampling to `spacing` along `axis`."""
        n_out = resample_size(self.shape[axis], self.resolution[axis], spacing)
        key = ("r", axis, n_out)
        p = self._ptrs.get(key)
        if p is None:
            t, fac = self.tables.resample(self.shape[axis], self.resolution[axis], spacing)
            p = self._ptrs[key] = (t.data_ptr(), fac)
        return p

    def zoom_back_ptr(self, axis: int, n_coarse: int, factor: float) -> int:
        """Table of the zoom from the coarse extent back to the volume extent (factor = the down-sampling's)."""
        key = ("b", axis, n_coarse)
        p = self._ptrs.get(key)
        if p is None:
            p = self._ptrs[key] = self.tables.zoom(n_coarse, float(1 / np.float64(factor)), self.shape[axis]).data_ptr()
        return p

    RING_SLOTS, RING_FLOATS = 8, 1 << 19

    # ------------------------------------------------------------------ launch batching
    def begin(self):
        """Open a launch batch: until ``flush`` the small-parameter uploads accumulate in ONE pinned
        ring slot and the C-ABI calls are queued, so a whole pipeline step costs one H2D copy
        followed by its launches (all job structs are built before the first kernel starts)."""
        if self._batch is not None:
            raise RuntimeError("SynthEngine.begin: a launch batch is already open")
        self._ring_init()
        hring, dring, events, pos = self._ring
        slot = pos[0] % self.RING_SLOTS
        pos[0] += 1
        if events[slot] is not None:
            events[slot].synchronize()
        self._batch = {"slot": slot, "used": 0, "calls": [], "keep": []}
        self.tables.hold = True

    def flush(self):
        b, self._batch = self._batch, None
        self.tables.hold = False
        hring, dring, events, _ = self._ring
        stream = _stream()
        if b["used"]:
            # SM fetch from the pinned (device-mapped) ring, not a copy-engine transfer: see fsg_fetch_params
            _lib.call("fsg_fetch_params", hring[b["slot"]].data_ptr(), dring[b["slot"]].data_ptr(), b["used"], stream)
            e = torch.cuda.Event()
            e.record()
            events[b["slot"]] = e
        for name, args in b["calls"]:
            _lib.call(name, *args, stream)
        self._keep_batch = b["keep"]

    def _call(self, name, *args):
        """C-ABI call on the current stream (queued while a launch batch is open)."""
        if self._batch is not None:
            self._batch["calls"].append((name, args))
        else:
            _lib.call(name, *args, _stream())

    def _ring_init(self):
        if not hasattr(self, "_ring"):
            self._ring = (torch.empty((self.RING_SLOTS, self.RING_FLOATS), dtype=torch.float32, pin_memory=True),
                          torch.empty((self.RING_SLOTS, self.RING_FLOATS), dtype=torch.float32, device=self.device),
                          [None] * self.RING_SLOTS, [0])
            self._ring_np = self._ring[0].numpy()

    def upload(self, arrays: list[np.ndarray]) -> list[torch.Tensor]:
        """One H2D copy for a list of small float32 host arrays; returns device views.  Staged
        through a ring of pinned slots with a matching device ring (no allocation per call)."""
        sizes = [int(a.size) for a in arrays]
        total = sum((s + 3) // 4 * 4 for s in sizes) or 4
        b = self._batch
        if b is not None and b["used"] + total <= self.RING_FLOATS:
            hv, dev = self._ring_np[b["slot"]], self._ring[1][b["slot"]]
            o, views = b["used"], []
            for a, s in zip(arrays, sizes):
                hv[o : o + s] = np.asarray(a, dtype=np.float32).reshape(-1)
                views.append(dev[o : o + s])
                o += (s + 3) // 4 * 4
            b["used"] = o
            return views
        if total > self.RING_FLOATS or b is not None:  # oversized request: one-off buffers
            host = torch.empty(total, dtype=torch.float32, pin_memory=True)
            dev = torch.empty(total, dtype=torch.float32, device=self.device)
            ev = None
        else:
            self._ring_init()
            hring, dring, events, pos = self._ring
            slot = pos[0] % self.RING_SLOTS
            pos[0] += 1
            if events[slot] is not None:
                events[slot].synchronize()  # the slot's previous copy and its consumers have run
            host, dev = hring[slot, :total], dring[slot, :total]
            ev = slot
        hv = host.numpy()
        o, views = 0, []
        for a, s in zip(arrays, sizes):
            hv[o : o + s] = np.asarray(a, dtype=np.float32).reshape(-1)
            views.append(dev[o : o + s])
            o += (s + 3) // 4 * 4
        _lib.call("fsg_fetch_params", host.data_ptr(), dev.data_ptr(), total, _stream())
        if ev is not None:
            # host slot: reusable once this copy has run; device slot: later copies into it are
            # stream-ordered after the kernels that read it (everything runs on the current stream)
            e = torch.cuda.Event()
            e.record()
            self._ring[2][ev] = e
        elif self._batch is not None:
            self._batch["keep"].append((host, dev))  # read by launches that are still queued
        else:
            self._oversize_keep = (host, dev)
        return views

    # ------------------------------------------------------------------ K1
    def gmm(self, plans, seeds, out: torch.Tensor, labels_out=None, pairs=None, tex=None):
        """seeds[b]: list of 1..4 int8/uint8 device volumes summed into the label map, or ``(PackedSeeds,
        mlabel2subclusters)``: the labels are decoded in the kernel from the subject's bit-packed seed words
        (2 bytes per voxel, no unpack pass).  pairs[b] True: sample b is written in the fixed-point pairs format
        of the warp fast path (same 4 bytes per voxel, into the same buffer) instead of float32.  tex[b] a
        ``TexVolume``: sample b is written into that block-linear volume instead of out[b]."""
        B = len(plans)
        nvox = int(out.shape[-1]) if torch.is_tensor(out) else self.nvox  # a list mixes [N] outputs and [2N] float-pair outputs
        small = self.upload([p.mus for p in plans] + [p.sigmas for p in plans])
        jobs = (_lib.GmmJob * B)()
        kinds = []
        for b, p in enumerate(plans):
            j = jobs[b]
            sd = seeds[b]
            if isinstance(sd, tuple):  # (PackedSeeds, mlabel2subclusters)
                ps, m2s = sd
                if int(np.prod(ps.shape)) != nvox:
                    raise ValueError("packed seed words must hold one word per output voxel")
                words = ps.on(self.device)
                for m in range(1, 5):
                    n = int(m2s[m])
                    if n not in ps.layout:
                        raise KeyError(f"no seeds with {n} sub-classes in this cache (available: {ps.counts})")
                    j.shift[m - 1], j.mask[m - 1] = ps.layout[n]
                j.words, j.word_bytes = words.data_ptr(), ps.word_bytes
                kinds.append(0)
            else:
                vols = list(sd)
                if not 1 <= len(vols) <= 4:
                    raise ValueError("each sample needs 1..4 seed volumes")
                for m, v in enumerate(vols):
                    if v.dtype not in (torch.int8, torch.uint8) or v.numel() != nvox:
                        raise TypeError("seed volumes must be int8/uint8 tensors with one label per output voxel")
                    _check(v, v.dtype, self.device, "seed volume")
                    j.seed[m] = v.data_ptr()
                kinds.append(len(vols))
            j.mus, j.sigmas = small[b].data_ptr(), small[B + b].data_ptr()
            j.nlabels = int(p.mus.size)
            j.noise = _ptr(None if p.gmm_noise is None else _check(p.gmm_noise, torch.float32, self.device, "gmm_noise"))
            if tex is not None and tex[b] is not None:
                j.out, j.out_surf, j.row_len, j.surf_ny = None, tex[b].h.surf, int(self.shape[2]), int(self.shape[1])
            elif pairs is not None and pairs[b]:
                j.out, j.out_pairs, j.row_len = None, out[b].data_ptr(), int(self.shape[2])
                j.pairs_float = 1 if _PAIRS_MODE == 2 else 0
            else:
                j.out = out[b].data_ptr()
            j.labels_out = None if labels_out is None else labels_out[b].data_ptr()
            j.rng = _lib.Rng(p.rng_seed & (2**64 - 1), p.sample_id, STAGE_GMM, 0)
        # one launch per kind of label source and noise source (a launch needs them uniform); production batches have one
        groups: dict = {}
        for b, k in enumerate(kinds):
            groups.setdefault((k, plans[b].gmm_noise is not None), []).append(b)
        if len(groups) == 1:
            self._call("fsg_gmm", jobs, B, nvox)
        else:
            for idx in groups.values():
                sub = (_lib.GmmJob * len(idx))(*[jobs[b] for b in idx])
                self._call("fsg_gmm", sub, len(idx), nvox)
        self._keep = (small,)

    # ------------------------------------------------------------------ K2
    def _warp_jobs(self, plans, src_img, src_seg, dst_img, dst_seg, src_img2=None, dst_img2=None, epilogue=True, pairs=None, tex=None):
        B = len(plans)
        sx, sy, sz = self.shape
        arrays, slots, gjobs = [], [], []
        for b, p in enumerate(plans):
            slot = {}
            if p.deform and p.fsmall is not None:
                slot["f"] = len(arrays)
                arrays.append(p.fsmall)
            elif p.deform and p.fsmall_dev is not None:
                gjobs.append((b, "f", p.fsmall_dev[0], 3 * int(np.prod(p.fsmall_dev[0])), p.fsmall_dev[1], STAGE_FIELD))
            if epilogue and p.bf_low is not None:
                slot["b"] = len(arrays)
                arrays.append(p.bf_low)
            elif epilogue and p.bf_dev is not None:
                gjobs.append((b, "b", p.bf_dev[0], int(np.prod(p.bf_dev[0])), p.bf_dev[1], STAGE_BIAS))
            slots.append(slot)
        small = self.upload(arrays) if arrays else []
        dev_grid = {}
        if gjobs:
            # control grids drawn on the device: one launch for the whole batch, no upload
            nf = (max([g[3] for g in gjobs if g[1] == "f"], default=0) + 3) // 4 * 4
            nb = (max([g[3] for g in gjobs if g[1] == "b"], default=0) + 3) // 4 * 4
            gbuf = self.scratch("grids", B, torch.float32, max(nf + nb, 4))
            gj = (_lib.GridJob * len(gjobs))()
            for q, (b, kind, shp, n, std, stage) in enumerate(gjobs):
                ptr = gbuf[b].data_ptr() + (0 if kind == "f" else 4 * nf)
                gj[q].out, gj[q].n, gj[q].scale = ptr, n, float(std)
                gj[q].rng = _lib.Rng(plans[b].rng_seed & (2**64 - 1), plans[b].sample_id, stage, 0)
                dev_grid[(b, kind)] = (ptr, shp)
            for q0 in range(0, len(gjobs), _lib.MAX_JOBS):
                chunk = (_lib.GridJob * min(_lib.MAX_JOBS, len(gjobs) - q0))(*gj[q0 : q0 + _lib.MAX_JOBS])
                self._call("fsg_draw_grids", chunk, len(chunk))
        shift = self.scratch("shift", B, torch.float32, 4)
        jobs = (_lib.WarpJob * B)()
        keep = [small, shift]
        for b, p in enumerate(plans):
            j = jobs[b]
            j.src_img, j.dst_img = _ptr(None if src_img is None else src_img[b]), _ptr(None if dst_img is None else dst_img[b])
            if tex is not None and tex[b] is not None:
                j.src_img, j.src_tex = None, tex[b].h.tex
            elif pairs is not None and pairs[b]:
                j.src_img, j.src_pairs = None, _ptr(src_img[b])
                j.pairs_float = 1 if _PAIRS_MODE == 2 else 0
            j.src_seg, j.dst_seg = _ptr(None if src_seg is None else src_seg[b]), _ptr(None if dst_seg is None else dst_seg[b])
            j.src_img2, j.dst_img2 = _ptr(None if src_img2 is None else src_img2[b]), _ptr(None if dst_img2 is None else dst_img2[b])
            j.mode, j.flip = int(bool(p.deform)), int(bool(p.flip))
            j.shift = shift[b].data_ptr()
            if p.deform:
                j.A = (C.c_float * 9)(*np.asarray(p.A, dtype=np.float32).reshape(-1))
                j.c2 = (C.c_float * 3)(*np.asarray(p.c2, dtype=np.float64).astype(np.float32))
                j.center = (C.c_float * 3)(*np.asarray(p.center, dtype=np.float32))
                if "f" in slots[b] or (b, "f") in dev_grid:
                    if "f" in slots[b]:
                        fs = p.fsmall.shape[:3]
                        j.fsmall = small[slots[b]["f"]].data_ptr()
                    else:
                        j.fsmall, fs = dev_grid[(b, "f")]
                    j.fs = (C.c_int32 * 3)(*fs)
                    for a in range(3):
                        j.ftab[a] = self.tables.zoom(fs[a], self.shape[a] / fs[a], self.shape[a]).data_ptr()
            if epilogue and p.gamma is not None:
                j.has_gamma, j.gamma = 1, float(np.float32(p.gamma))
            if "b" in slots[b] or (b, "b") in dev_grid:
                if "b" in slots[b]:
                    bs = p.bf_low.shape
                    j.bf_low = small[slots[b]["b"]].data_ptr()
                else:
                    j.bf_low, bs = dev_grid[(b, "b")]
                j.bs = (C.c_int32 * 3)(*bs)
                for a in range(3):
                    j.btab[a] = self.tables.zoom(bs[a], self.shape[a] / bs[a], self.shape[a]).data_ptr()
        return jobs, keep

    def warp(self, plans, src_img, src_seg, dst_img, dst_seg, src_img2=None, dst_img2=None, epilogue=True, pairs=None, tex=None):
        B = len(plans)
        sx, sy, sz = self.shape
        jobs, keep = self._warp_jobs(plans, src_img, src_seg, dst_img, dst_seg, src_img2, dst_img2, epilogue, pairs, tex)
        didx = [b for b, p in enumerate(plans) if p.deform]
        if didx:
            dj = (_lib.WarpJob * len(didx))(*[jobs[b] for b in didx])
            self._call("fsg_warp_shift", dj, len(didx), sx, sy, sz)
        self._call("fsg_warp", jobs, B, sx, sy, sz)
        self._keep_warp = keep

    def warp_coords(self, plan):
        sx, sy, sz = self.shape
        jobs, keep = self._warp_jobs([plan], None, None, None, None, epilogue=False)
        self._call("fsg_warp_shift", jobs, 1, sx, sy, sz)
        out = torch.empty((3, sx, sy, sz), dtype=torch.float32, device=self.device)
        self._call("fsg_warp_coords", jobs, sx, sy, sz, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr())
        return out

    # ------------------------------------------------------------------ K4a
    def blur(self, stds_list, src, dst, tmp):
        """Separable Gaussian blur of src[b] into dst[b] (tmp[b] = scratch); stds per sample."""
        B = len(stds_list)
        sx, sy, sz = self.shape
        jobs = (_lib.BlurJob * B)()
        keep = []
        for b, stds in enumerate(stds_list):
            j = jobs[b]
            j.src, j.dst, j.tmp = src[b].data_ptr(), dst[b].data_ptr(), tmp[b].data_ptr()
            for a in range(3):
                if stds[a] > 0:
                    t = self.tables.taps(float(stds[a]))
                    keep.append(t)
                    j.taps[a], j.ntaps[a] = t.data_ptr(), t.numel() // 4
        self._call("fsg_blur3d", jobs, B, sx, sy, sz)

    # ------------------------------------------------------------------ K4ab (fused)
    def sep_extents(self, p, positions=True):
        """Coarse extents of one sample's blur + down-sampling (``positions=False``: plain blur)."""
        if not positions:
            return tuple(self.shape)
        return tuple(resample_size(self.shape[a], self.resolution[a], p.spacing[a]) for a in range(3))

    def sep_capacity(self, p, positions=True):
        """Floats ``sepconv`` needs behind (dst, tmp1, tmp2) for this sample.  They exceed the volume when an
        axis is up-sampled (simulated spacing finer than the input resolution, synthseg.py:80-84)."""
        n = self.sep_extents(p, positions)
        sx, sy, sz = self.shape
        return n[0] * n[1] * n[2], n[0] * sy * sz, n[0] * n[1] * sz

    def sepconv(self, plans, src, dst, tmp1, tmp2, positions=True):
        """Fused blur + trilinear down-sampling (+noise): plan.stds gives the blur, plan.spacing the
        coarse grid (positions=False: plain blur at full resolution).  tmp1 may alias dst.
        Returns per-sample (coarse shape, factors) like ``resample``.  Every buffer must hold what
        ``sep_capacity`` reports for its sample (checked here and again by the library)."""
        B = len(plans)
        sx, sy, sz = self.shape
        tap_arrays, tap_slot, maxw = [], [], 2
        for p in plans:
            slots, seen = [], {}
            for a in range(3):
                sd = float(p.stds[a]) if p.stds is not None else 0.0
                if sd > 0:
                    if sd not in seen:  # isotropic spacing gives three identical widths: upload once
                        seen[sd] = len(tap_arrays)
                        tap_arrays.append(gaussian_taps_np(sd))
                    slots.append(seen[sd])
                    maxw = max(maxw, tap_arrays[seen[sd]].size + 1)
                else:
                    slots.append(None)
            tap_slot.append(slots)
        taps_dev = self.upload(tap_arrays) if tap_arrays else []
        extents = [self.sep_extents(p, positions) for p in plans]
        for b, n in enumerate(extents):
            need = (n[0] * n[1] * n[2], n[0] * sy * sz, n[0] * n[1] * sz)
            for name, t, cnt in (("dst", dst[b], need[0]), ("tmp1", tmp1[b], need[1]), ("tmp2", tmp2[b], need[2])):
                if t.shape[-1] < cnt:
                    raise ValueError(f"sepconv: sample {b}: {name} holds {t.shape[-1]} floats, {cnt} are needed (coarse grid {n})")
        # workspace per job and axis: float w[nmax][maxw] then int16 q0[nmax]; an up-sampled axis has more rows than the volume
        nmax = max(max(self.shape), max(max(n) for n in extents))
        maxw = max(32, (maxw + 3) // 4 * 4)
        per_axis = (nmax * maxw + (nmax + 1) // 2 + 3) // 4 * 4
        ws = self.scratch("sep_tables", B, torch.float32, 3 * per_axis)
        cjobs = (_lib.SepComposeJob * (3 * B))()
        jobs = (_lib.SepconvJob * B)()
        info = []
        for b, p in enumerate(plans):
            j = jobs[b]
            n, factors = [], []
            for a in range(3):
                cj = cjobs[3 * b + a]
                n_in = self.shape[a]
                if positions:
                    t, fac = self.tables.resample(n_in, self.resolution[a], p.spacing[a])
                    n_out = extents[b][a]
                    cj.pos = t.data_ptr()
                else:
                    n_out, fac = n_in, 1.0
                ntaps = 1
                if tap_slot[b][a] is not None:
                    td = taps_dev[tap_slot[b][a]]
                    cj.taps, ntaps = td.data_ptr(), int(td.numel())
                width = min(n_in, ntaps + (1 if positions else 0))
                w_ptr = ws[b].data_ptr() + 4 * a * per_axis
                q_ptr = w_ptr + 4 * nmax * maxw
                cj.q0_out, cj.w_out = q_ptr, w_ptr
                cj.ntaps, cj.n_in, cj.n_out, cj.width = ntaps, n_in, n_out, width
                cj.cap_q0, cj.cap_w = nmax, nmax * maxw
                ax = j.ax[a]
                ax.q0, ax.w, ax.n_out, ax.width = q_ptr, w_ptr, n_out, width
                ax.pos, ax.taps, ax.ntaps = cj.pos, cj.taps, ntaps
                n.append(n_out)
                factors.append(fac)
            j.src, j.dst, j.tmp1, j.tmp2 = src[b].data_ptr(), dst[b].data_ptr(), tmp1[b].data_ptr(), tmp2[b].data_ptr()
            j.cap_dst, j.cap_tmp1, j.cap_tmp2 = dst[b].numel(), tmp1[b].numel(), tmp2[b].numel()
            if p.noise_std is not None and positions:
                j.has_noise, j.noise_std = 1, float(np.float32(p.noise_std))
                j.noise = _ptr(None if p.noise is None else _check(p.noise, torch.float32, self.device, "noise"))
                j.rng = _lib.Rng(p.rng_seed & (2**64 - 1), p.sample_id, STAGE_NOISE, 0)
            info.append((tuple(n), np.asarray(factors, dtype=np.float64)))
        self._call("fsg_sep_compose", cjobs, 3 * B)
        self._call("fsg_sepconv", jobs, B, sx, sy, sz)
        self._keep_sep = taps_dev
        return info

    # ------------------------------------------------------------------ K4b
    def lowres_shape(self, spacing):
        return tuple(resample_size(self.shape[a], self.resolution[a], spacing[a]) for a in range(3))

    def resample(self, plans, src, dst):
        """Trilinear down-sampling (+noise when plan.noise_std is set). Returns per-sample (shape, factors)."""
        B = len(plans)
        sx, sy, sz = self.shape
        jobs = (_lib.ResampleJob * B)()
        info = []
        for b, p in enumerate(plans):
            j = jobs[b]
            n = self.lowres_shape(p.spacing)
            factors = []
            for a in range(3):
                t, fac = self.tables.resample(self.shape[a], self.resolution[a], p.spacing[a])
                j.tab[a] = t.data_ptr()
                factors.append(fac)
            j.n = (C.c_int32 * 3)(*n)
            j.src, j.dst = src[b].data_ptr(), dst[b].data_ptr()
            if p.noise_std is not None:
                j.has_noise, j.noise_std = 1, float(np.float32(p.noise_std))
                j.noise = _ptr(None if p.noise is None else _check(p.noise, torch.float32, self.device, "noise"))
                j.rng = _lib.Rng(p.rng_seed & (2**64 - 1), p.sample_id, STAGE_NOISE, 0)
            info.append((n, np.asarray(factors, dtype=np.float64)))
        self._call("fsg_resample", jobs, B, sx, sy, sz)
        return info

    def add_noise(self, plans, src, dst, numel=None):
        B = len(plans)
        jobs = (_lib.NoiseJob * B)()
        for b, p in enumerate(plans):
            j = jobs[b]
            j.src, j.dst = src[b].data_ptr(), dst[b].data_ptr()
            j.noise_std = float(np.float32(p.noise_std))
            j.noise = _ptr(None if p.noise is None else _check(p.noise, torch.float32, self.device, "noise"))
            j.rng = _lib.Rng(p.rng_seed & (2**64 - 1), p.sample_id, STAGE_NOISE, 0)
        self._call("fsg_add_noise", jobs, B, self.nvox if numel is None else numel)

    # ------------------------------------------------------------------ K4c
    def zoom(self, src_list, src_shapes, factors_list, dst, post=0, minmax=None):
        """myzoom_torch(src, factors) into dst[b] of the engine shape; post 1 = /max, 2 = /max + ScaleIntensity."""
        B = len(src_list)
        sx, sy, sz = self.shape
        jobs = (_lib.ZoomJob * B)()
        mm = self.scratch("minmax", B, torch.float32, 2) if minmax is None else minmax
        for b in range(B):
            j = jobs[b]
            n = src_shapes[b]
            for a in range(3):
                j.tab[a] = self.tables.zoom(n[a], factors_list[b][a], self.shape[a]).data_ptr()
            j.n = (C.c_int32 * 3)(*n)
            j.src, j.dst = src_list[b].data_ptr(), dst[b].data_ptr()
            j.minmax, j.post = mm[b].data_ptr(), post
        if post > 0:
            self._call("fsg_zoom_minmax", jobs, B, sx, sy, sz)
        self._call("fsg_zoom", jobs, B, sx, sy, sz)
        return mm

    # ------------------------------------------------------------------ misc
    def scale_intensity(self, x: torch.Tensor, out: torch.Tensor | None = None):
        out = torch.empty_like(x) if out is None else out
        mm = torch.empty(2, dtype=torch.float32, device=self.device)
        if self._batch is not None:
            self._batch["keep"].append(mm)
        self._call("fsg_minmax", x.data_ptr(), x.numel(), mm.data_ptr())
        self._call("fsg_scale_intensity", x.data_ptr(), out.data_ptr(), x.numel(), mm.data_ptr())
        return out

    def to_u8(self, x: torch.Tensor) -> torch.Tensor:
        if x.dtype == torch.uint8:
            return x.contiguous()
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        out = torch.empty(x.shape, dtype=torch.uint8, device=self.device)
        self._call("fsg_f32_to_u8", x.data_ptr(), out.data_ptr(), x.numel())
        return out

    def from_u8(self, x: torch.Tensor, dtype) -> torch.Tensor:
        if dtype == torch.uint8:
            return x
        out = torch.empty(x.shape, dtype=dtype, device=self.device)
        if dtype == torch.float32:
            self._call("fsg_u8_to_f32", x.data_ptr(), out.data_ptr(), x.numel())
        elif dtype == torch.int64:
            self._call("fsg_u8_to_i64", x.data_ptr(), out.data_ptr(), x.numel())
        else:
            out = self.from_u8(x, torch.float32).to(dtype)
        return out

    # ------------------------------------------------------------------ fused base pipeline
    def run_base(self, plans, seeds, segs, out_img=None, out_seg=None, scale=False):
        """Whole base path for a batch.  seeds[b] = 1..4 label volumes, segs[b] = uint8 volume.
        Returns (images [B,*shape] float32, segs [B,*shape] uint8)."""
        B = len(plans)
        if B < 1 or B > _lib.MAX_JOBS:
            raise ValueError(f"batch must be 1..{_lib.MAX_JOBS}")
        if len(seeds) != B or len(segs) != B:
            raise ValueError("run_base: plans, seeds and segmentations must have the same length")
        for b, sg in enumerate(segs):
            if sg.numel() != self.nvox:
                raise ValueError(f"segmentation {b} has {sg.numel()} voxels, the engine shape {self.shape} has {self.nvox}")
            _check(sg, torch.uint8, self.device, f"segmentation {b}")
        shp = (B, *self.shape)
        for name, t, dt in (("out_img", out_img, torch.float32), ("out_seg", out_seg, torch.uint8)):
            if t is not None:
                if tuple(t.shape) != shp:
                    raise ValueError(f"{name} must have shape {shp}, got {tuple(t.shape)}")
                _check(t, dt, self.device, name)
        out_img = torch.empty(shp, dtype=torch.float32, device=self.device) if out_img is None else out_img
        out_seg = torch.empty(shp, dtype=torch.uint8, device=self.device) if out_seg is None else out_seg
        buf0 = self.scratch("buf0", B)
        buf1 = self.scratch("buf1", B)
        buf2 = self.scratch("buf2", B)
        self.begin()
        try:
            self._run_base(plans, seeds, segs, out_img, out_seg, scale, buf0, buf1, buf2)
        except BaseException:
            self._batch = None
            self.tables.hold = False
            raise
        self.flush()
        return out_img, out_seg

    def pairs_eligible(self, p) -> bool:
        """The GMM image of this sample is only ever read by the warp fast path (deformation with a
        control grid, tile-aligned extents): hand it over as 16-bit fixed-point z-pairs."""
        sx, sy, sz = self.shape
        return bool(_PAIRS and p.deform and (p.fsmall is not None or p.fsmall_dev is not None) and sx % 8 == 0 and sy % 4 == 0 and sz % 4 == 0 and min(self.shape) >= 2)

    def tex_eligible(self, p) -> bool:
        """Same condition as ``pairs_eligible`` (the GMM image is read by the warp fast path only): hand it over
        in a block-linear texture volume."""
        sx, sy, sz = self.shape
        return bool(self.use_tex and p.deform and (p.fsmall is not None or p.fsmall_dev is not None) and sx % 8 == 0 and sy % 4 == 0 and sz % 4 == 0 and min(self.shape) >= 2)

    def _run_base(self, plans, seeds, segs, out_img, out_seg, scale, buf0, buf1, buf2):
        B = len(plans)
        tex = [self.texvol(b) if self.tex_eligible(p) else None for b, p in enumerate(plans)]
        pairs = [tex[b] is None and self.pairs_eligible(p) for b, p in enumerate(plans)]
        gmm_out = buf0
        if _PAIRS_MODE == 2 and any(pairs):  # float2 pairs need 8 bytes per voxel: their own scratch volume
            wide = self.scratch("gmm_fpairs", B, torch.float32, 2 * self.nvox)
            gmm_out = [wide[b] if pairs[b] else buf0[b] for b in range(B)]
        self.gmm(plans, seeds, gmm_out, pairs=pairs, tex=tex)
        rs = [b for b, p in enumerate(plans) if p.spacing is not None]
        no_rs = [b for b, p in enumerate(plans) if p.spacing is None]
        # warp straight into the output for samples that skip the resolution simulation
        warp_dst = [out_img[b].view(-1) if (plans[b].spacing is None and plans[b].noise_std is None) else buf1[b] for b in range(B)]
        self.warp(plans, gmm_out, segs, warp_dst, out_seg, pairs=pairs, tex=tex)
        if rs:
            sub = [plans[b] for b in rs]
            # x pass -> buf2, y pass -> buf0 (the GMM image is dead), z pass (+noise) -> buf2
            low, tmp = [buf2[b] for b in rs], [buf0[b] for b in rs]
            need = None
            if any(p.spacing[a] < self.resolution[a] for p in sub for a in range(3)):
                need = [self.sep_capacity(p) for p in sub]
            if need is not None and max(max(nd) for nd in need) > self.nvox:  # an up-sampled axis (spacing < resolution): the coarse grid outgrows the volume
                big_a = self.scratch("sep_big_a", len(rs), numel=max(max(nd[0], nd[1]) for nd in need))
                big_b = self.scratch("sep_big_b", len(rs), numel=max(nd[2] for nd in need))
                low, tmp = [big_a[k] for k in range(len(rs))], [big_b[k] for k in range(len(rs))]
            info = self.sepconv(sub, [buf1[b] for b in rs], low, low, tmp)
            self.zoom(low, [i[0] for i in info], [1 / i[1] for i in info], [out_img[b].view(-1) for b in rs], post=2 if scale else 1)
        nz = [b for b in no_rs if plans[b].noise_std is not None]
        if nz:
            self.add_noise([plans[b] for b in nz], [buf1[b] for b in nz], [out_img[b].view(-1) for b in nz])
        if scale and no_rs:
            for b in no_rs:
                self.scale_intensity(out_img[b], out_img[b])
        return out_img, out_seg


_ENGINES: dict = {}


def engine_for(device, shape, resolution=(1.0, 1.0, 1.0)) -> SynthEngine:
    """Process-wide engine cache keyed by (device, shape, resolution, current CUDA stream).  An engine's scratch
    volumes, workspace and parameter ring are ordered by stream order alone, so every stream that generates
    (a ``DeviceBatchLoader``'s producer stream next to the consumer's default stream, a second loader, ...) gets
    its own engine: nothing is shared across streams, no cross-stream events are needed."""
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
    stream = torch.cuda.current_stream(dev).cuda_stream if (dev.type == "cuda" and torch.cuda.is_available()) else 0
    key = (str(dev), tuple(int(s) for s in shape), tuple(float(r) for r in resolution), int(stream))
    eng = _ENGINES.get(key)
    if eng is None:
        eng = SynthEngine(shape, resolution, dev)
        _ENGINES[key] = eng
    return eng
