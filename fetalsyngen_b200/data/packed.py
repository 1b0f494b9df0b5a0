"""Bit-packed per-subject seed cache.

The reference stores a subject's seeds as one int8 NIfTI per (sub-class count n, meta-label m) —
24 gzip files for counts 1..6 — and decodes four of them per sample (``rand_gmm.py:90-97``,
0.23 s on the host).  All of those volumes share the meta-label support, so a single word per voxel
carries them: bits 0-2 the meta-label (0 background, 1..4), then for every count n >= 2 a field of
ceil(log2 n) bits with the voxel's sub-class index.  Counts 1..6 need 14 bits (uint16, 32 MiB per
256^3 subject instead of 384 MiB of int8 volumes), counts 1..10 need 28 (uint32).

``pack_seed_volumes`` is the one-time converter, ``PackedSeeds`` the device-resident form: the label
volume a sample needs (what summing the four selected seed files gives) is produced by
``fsg_unpack_seeds`` from the counts drawn for the sample.  ``save_packed`` / ``load_packed`` keep the
words together with the uint8 segmentation in one uncompressed ``.npz`` per subject, so a process
start costs one file read per subject instead of 25 gzip decodes.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from .. import _lib

FORMAT_VERSION = 1


def field_layout(counts) -> dict:
    """{n: (shift, mask)} of the sub-class fields for the given sub-class counts (ascending)."""
    counts = sorted(int(c) for c in counts)
    if not counts or counts[0] < 1:
        raise ValueError("sub-class counts must be positive")
    layout, shift = {}, 3
    for n in counts:
        bits = int(np.ceil(np.log2(n))) if n > 1 else 0
        if bits > 4:
            raise ValueError(f"{n} sub-classes do not fit a 4-bit field")
        layout[n] = (shift if bits else 0, (1 << bits) - 1)
        shift += bits
    if shift > 32:
        raise ValueError("the sub-class fields do not fit a 32-bit word")
    return layout


def word_dtype(layout: dict):
    top = max((s + int(m).bit_length() for s, m in layout.values()), default=3)
    return np.uint16 if max(top, 3) <= 16 else np.uint32


def pack_seed_volumes(seeds: dict) -> tuple[np.ndarray, list]:
    """``seeds[n][m]`` = int8 volume of meta-label m (1..4) split into n sub-classes, values 0 or
    ``10 m + k`` (scripts/generate_seeds.py:207-209) -> (words, sorted counts).  Raises ValueError when
    the volumes cannot be packed losslessly (overlapping supports, supports that change with n, labels
    outside ``10 m .. 10 m + n - 1``)."""
    counts = sorted(int(n) for n in seeds)
    layout = field_layout(counts)
    dt = word_dtype(layout)
    meta = None
    words = None
    for n in counts:
        per = seeds[n]
        if sorted(int(m) for m in per) != [1, 2, 3, 4]:
            raise ValueError(f"sub-class count {n}: expected meta-labels 1..4, got {sorted(per)}")
        meta_n = None
        sub = None
        for m in range(1, 5):
            v = np.asarray(per[m])
            on = v != 0
            if meta_n is None:
                meta_n = np.zeros(v.shape, dtype=np.uint8)
                sub = np.zeros(v.shape, dtype=np.uint8)
            if (meta_n[on] != 0).any():
                raise ValueError(f"sub-class count {n}: the supports of the meta-labels overlap")
            k = v[on].astype(np.int32) - 10 * m
            if k.size and (k.min() < 0 or k.max() >= n):
                raise ValueError(f"sub-class count {n}, meta-label {m}: labels outside {10 * m}..{10 * m + n - 1}")
            meta_n[on] = m
            sub[on] = k
        if meta is None:
            meta = meta_n
            words = meta.astype(dt)
        elif not np.array_equal(meta, meta_n):
            raise ValueError(f"sub-class count {n}: meta-label support differs from count {counts[0]}")
        shift, _ = layout[n]
        words |= sub.astype(dt) << dt(shift)
    return np.ascontiguousarray(words), counts


def unpack_numpy(words: np.ndarray, counts, mlabel2subclusters: dict) -> np.ndarray:
    """Host restatement of ``fsg_unpack_seeds`` (used by the converter's self-check and the CPU tests)."""
    layout = field_layout(counts)
    meta = (words & 7).astype(np.int32)
    out = np.zeros(words.shape, dtype=np.uint8)
    for m in range(1, 5):
        shift, mask = layout[int(mlabel2subclusters[m])]
        sel = meta == m
        out[sel] = (10 * m + ((words[sel].astype(np.int64) >> shift) & mask)).astype(np.uint8)
    return out


class PackedSeeds:
    """A subject's packed seed words on a device.  ``labels(mlabel2subclusters)`` returns the uint8
    label volume for one draw of sub-class counts (one kernel, 2-4 bytes read and 1 written per voxel)."""

    def __init__(self, words, counts, device=None):
        w = np.ascontiguousarray(words)
        if w.dtype not in (np.uint16, np.uint32):
            raise ValueError("packed words must be uint16 or uint32")
        self.counts = [int(c) for c in counts]
        self.layout = field_layout(self.counts)
        if word_dtype(self.layout) != w.dtype:
            raise ValueError(f"word type {w.dtype} does not match the layout of counts {self.counts}")
        self.shape = tuple(w.shape)
        self.word_bytes = w.dtype.itemsize
        self._host = w
        self._dev: dict = {}
        if device is not None:
            self.on(device)

    @classmethod
    def from_device(cls, words: torch.Tensor, counts) -> "PackedSeeds":
        """Wrap packed words that already live on a device (int16 / int32 tensor holding the uint16 / uint32
        bit patterns), e.g. a slot of the host pipeline that has just been uploaded."""
        if words.dtype not in (torch.int16, torch.int32) or words.device.type != "cuda":
            raise ValueError("from_device expects an int16 / int32 CUDA tensor")
        self = cls.__new__(cls)
        self.counts = [int(c) for c in counts]
        self.layout = field_layout(self.counts)
        self.word_bytes = words.element_size()
        if np.dtype(word_dtype(self.layout)).itemsize != self.word_bytes:
            raise ValueError(f"{8 * self.word_bytes}-bit words do not match the layout of counts {self.counts}")
        self.shape = tuple(words.shape)
        self._host = None
        self._dev = {str(words.device): words}
        return self

    def on(self, device) -> torch.Tensor:
        key = str(torch.device(device))
        t = self._dev.get(key)
        if t is None:
            signed = self._host.view(np.int16 if self.word_bytes == 2 else np.int32)  # bytes only: torch lacks most uint16/32 ops
            t = torch.from_numpy(signed).to(device)
            self._dev[key] = t
        return t

    def job(self, j, mlabel2subclusters: dict, device, out: torch.Tensor):
        for m in range(1, 5):
            n = int(mlabel2subclusters[m])
            if n not in self.layout:
                raise KeyError(f"no seeds with {n} sub-classes in this cache (available: {self.counts})")
            j.shift[m - 1], j.mask[m - 1] = self.layout[n]
        j.words, j.out, j.word_bytes = self.on(device).data_ptr(), out.data_ptr(), self.word_bytes

    def labels(self, mlabel2subclusters: dict, device) -> torch.Tensor:
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.FsgError("PackedSeeds.labels needs a CUDA device: libfsg has no CPU fallback")
        out = torch.empty(self.shape, dtype=torch.uint8, device=dev)
        jobs = (_lib.UnpackJob * 1)()
        self.job(jobs[0], mlabel2subclusters, dev, out)
        with torch.cuda.device(dev):
            _lib.call("fsg_unpack_seeds", jobs, 1, out.numel(), torch.cuda.current_stream().cuda_stream)
        return out


def save_packed(path, seg_u8: np.ndarray, words: np.ndarray, counts, affine=None) -> None:
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    tmp = path.with_suffix(".tmp.npz")
    np.savez(tmp, version=np.int32(FORMAT_VERSION), seg=np.ascontiguousarray(seg_u8, dtype=np.uint8), words=words, counts=np.asarray(counts, dtype=np.int32),
             affine=np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64))
    tmp.replace(path)  # a reader never sees a half-written cache file


def load_packed(path):
    """-> (uint8 segmentation, PackedSeeds, affine)."""
    with np.load(path) as z:
        if int(z["version"]) != FORMAT_VERSION:
            raise ValueError(f"{path}: packed cache version {int(z['version'])}, expected {FORMAT_VERSION}")
        return z["seg"], PackedSeeds(z["words"], z["counts"].tolist()), z["affine"]


def pack_subject(seg_path, seed_paths: dict, out_file, verify: bool = True):
    """One-time converter for one subject: ``seed_paths[n][m]`` NIfTI paths + the segmentation ->
    ``out_file`` (.npz).  ``verify`` re-derives every seed volume from the packed words before writing."""
    from ..utils.nifti import read_nifti, to_ras

    seg, affine = to_ras(*read_nifti(seg_path, with_affine=True))  # the cache holds RAS-oriented volumes, like the reference feeds its generator
    seg_f = np.nan_to_num(np.asarray(seg, dtype=np.float32))
    if seg_f.min() < 0 or seg_f.max() > 255 or not np.array_equal(seg_f, np.round(seg_f)):
        raise ValueError(f"{seg_path}: labels are not integers in 0..255")
    vols = {int(n): {int(m): to_ras(*read_nifti(p, with_affine=True))[0] for m, p in per.items()} for n, per in seed_paths.items()}
    words, counts = pack_seed_volumes(vols)
    if words.shape != seg_f.shape:
        raise ValueError(f"{seg_path}: segmentation {seg_f.shape} and seeds {words.shape} differ in shape")
    if verify:
        for n in counts:
            lab = unpack_numpy(words, counts, {m: n for m in range(1, 5)})
            want = sum(np.asarray(vols[n][m]).astype(np.int32) for m in range(1, 5)).astype(np.uint8)
            if not np.array_equal(lab, want):
                raise ValueError(f"{out_file}: packed words do not reproduce the seeds of count {n}")
    save_packed(out_file, seg_f.astype(np.uint8), words, counts, affine)
    return Path(out_file)
