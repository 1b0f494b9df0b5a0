"""Device-side building blocks of the SR artifacts: thin wrappers that turn host parameters into
calls of the K5 entry points of libfsg (``csrc/artifacts.cu``).  torch is used for memory only."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .tables import TAB_DTYPE

STAGE_SAMPLE, STAGE_RING, STAGE_PYRAMID = 16, 17, 32


def _stream():
    return torch.cuda.current_stream().cuda_stream


def upload(arr, dtype, device) -> torch.Tensor:
    """Small host array -> device without stalling the host: a blocking ``.to(device)`` from pageable
    memory synchronises the stream, i.e. waits for every kernel queued so far (measured: 0.7 ms per
    call inside SimulateMotion, ~40 calls per sample).  The copy goes through torch's caching pinned
    allocator instead, which keeps the staging block alive until the copy has run."""
    h = torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype))
    if torch.device(device).type != "cuda":
        return h.to(device)
    return h.pin_memory().to(device, non_blocking=True)


def upsample_table(n_in: int, n_out: int) -> np.ndarray:
    """1-D table of ``F.interpolate(mode='trilinear', align_corners=False)``
    (reference call site: augmentation/artifacts.py:315-320)."""
    scale = n_in / n_out
    src = np.maximum(scale * (np.arange(n_out, dtype=np.float64) + 0.5) - 0.5, 0).astype(np.float32)
    fl = np.floor(src)
    tab = np.zeros(n_out, dtype=TAB_DTYPE)
    tab["f"] = fl.astype(np.int16)
    tab["c"] = np.minimum(fl.astype(np.int32) + 1, n_in - 1).astype(np.int16)
    tab["wc"] = src - fl
    return tab


class ArtifactOps:
    def __init__(self, engine):
        self.eng = engine
        self.dev = engine.device
        self.shape = engine.shape
        self.n = engine.nvox

    # ------------------------------------------------------------------ buffers
    def u8(self, name):
        return self.eng.scratch("art_" + name, 1, torch.uint8)[0]

    def u16(self, name):
        return self.eng.scratch("art_" + name, 1, torch.int16)[0]

    def f32(self, name):
        return self.eng.scratch("art_" + name, 1, torch.float32)[0]

    # ------------------------------------------------------------------ MoG / sampling
    def mog(self, centers_axis, sigmas_axis, out=None, blend=None):
        """centers/sigmas: [n,3] float32 device tensors in axis order.  blend=(a, b, dst)."""
        n = int(centers_axis.shape[0])
        sx, sy, sz = self.shape
        a = b = dst = None
        if blend is not None:
            a, b, dst = (t.data_ptr() for t in blend)
        _lib.call("fsg_mog", centers_axis.data_ptr() if n else None, sigmas_axis.data_ptr() if n else None, n, sx, sy, sz,
                  None if out is None else out.data_ptr(), a, b, dst, _stream())

    def sample_voxels(self, labels, k, match=-1, labels2=None, prior=None, transpose_out=True, rng=(0, 0)):
        """Draw k voxel centres without replacement; returns ([k,3] float32 device tensor, count)."""
        sx, sy, sz = self.shape
        out = torch.zeros((max(k, 1), 3), dtype=torch.float32, device=self.dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=self.dev)
        ws = self.eng.scratch("art_sample_ws", 1, torch.uint8, 64 + 2048 * 8 + 64)[0]
        job = _lib.SampleJob()
        job.labels, job.labels2 = labels.data_ptr(), None if labels2 is None else labels2.data_ptr()
        job.centers_out, job.count_out = out.data_ptr(), cnt.data_ptr()
        job.workspace, job.workspace_bytes = ws.data_ptr(), ws.numel()
        job.rng = _lib.Rng(rng[0] & (2**64 - 1), rng[1], STAGE_SAMPLE, 0)
        job.match, job.k, job._pad = int(match), int(k), 1 if transpose_out else 0
        if prior is not None:
            pc, ps = prior
            job.nprior = len(pc)
            for q in range(len(pc)):
                for a in range(3):
                    job.prior_centers[q][a] = float(pc[q][a])
                    job.prior_sigmas[q][a] = float(ps[q][a])
        _lib.call("fsg_sample_voxels", C.byref(job), sx, sy, sz, _stream())
        return out, int(cnt.item())

    # ------------------------------------------------------------------ zoom with explicit tables
    def zoom_tabs(self, src, n_in, tabs, n_out, dst=None, post=0, minmax=None, reduce_only=False):
        job = (_lib.ZoomJob * 1)()
        j = job[0]
        j.src = src.data_ptr()
        j.dst = None if dst is None else dst.data_ptr()
        for a in range(3):
            j.tab[a] = tabs[a].data_ptr()
        j.n = (C.c_int32 * 3)(*n_in)
        j.post = post
        j.minmax = None if minmax is None else minmax.data_ptr()
        if reduce_only:
            _lib.call("fsg_zoom_minmax", job, 1, *n_out, _stream())
        else:
            _lib.call("fsg_zoom", job, 1, *n_out, _stream())

    def upsample_tabs(self, n_in, n_out):
        return [self.eng.tables._put(("up", n_in[a], n_out[a]), lambda a=a: upsample_table(n_in[a], n_out[a])) for a in range(3)]

    def add_noise_noclamp(self, buf, numel, rng, noise=None):
        """buf += N(0,1) (no clamp): one level of the StructNoise pyramid."""
        job = (_lib.NoiseJob * 1)()
        j = job[0]
        j.src = j.dst = buf.data_ptr()
        j.noise_std, j.flags = 1.0, 1
        j.noise = None if noise is None else noise.data_ptr()
        j.rng = rng
        _lib.call("fsg_add_noise", job, 1, int(numel), _stream())

    # ------------------------------------------------------------------ morphology
    def box(self, src, dst, tmp, k, op):
        _lib.call("fsg_morph_box", src.data_ptr(), dst.data_ptr(), tmp.data_ptr(), int(k), int(op), *self.shape, _stream())
        return dst

    def dist(self, mask, out16, tmp16, r, metric):
        _lib.call("fsg_morph_dist", mask.data_ptr(), out16.data_ptr(), tmp16.data_ptr(), int(r), int(metric), *self.shape, _stream())
        return out16

    def thresh(self, d16, out, thr):
        _lib.call("fsg_morph_thresh", d16.data_ptr(), out.data_ptr(), int(thr), self.n, _stream())
        return out
