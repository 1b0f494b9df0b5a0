"""Host-buffer front end of the batched generator: what a data-loading process calls when the
inputs (segmentation + seed volumes) and the outputs (image + segmentation) live in host
memory, as in the reference's ``FetalSynthDataset.sample`` (``fetalsyngen/data/datasets.py:
256-327``: host tensors in, ``.cpu()`` tensors out).

Software pipeline over ``depth`` buffer slots: pinned host input -> H2D (copy stream) ->
batched generation (compute stream) -> D2H (second copy stream) -> pinned host output.  The
three stages of consecutive steps overlap (PCIe is full duplex and the copy engines run beside
the SMs); ordering is by CUDA events only, the host blocks in ``collect`` alone.  Per step the
link carries 80 MiB per volume in each direction (uint8 segmentation + four int8 seed volumes
in, float32 image + uint8 segmentation out), so the pipeline is PCIe-bound, not kernel-bound.
"""
from __future__ import annotations

from collections import deque

import numpy as np
import torch


class _Slot:
    def __init__(self, batch, shape, dev, word_dtype=None, inputs: bool = True):
        shp = (batch, *shape)
        if inputs:
            self.h_seg = torch.empty(shp, dtype=torch.uint8, pin_memory=True)
            if word_dtype is None:
                self.h_seeds = torch.empty((batch, 4, *shape), dtype=torch.int8, pin_memory=True)
            else:  # bit-packed seed words of the subject cache (data/packed.py): one word per voxel instead of 4 bytes
                self.h_seeds = torch.empty(shp, dtype=word_dtype, pin_memory=True)
            self.d_seg = torch.empty(shp, dtype=torch.uint8, device=dev)
            self.d_seeds = torch.empty(self.h_seeds.shape, dtype=self.h_seeds.dtype, device=dev)
        self.h_img = torch.empty(shp, dtype=torch.float32, pin_memory=True)
        self.h_oseg = torch.empty(shp, dtype=torch.uint8, pin_memory=True)
        self.d_img = torch.empty(shp, dtype=torch.float32, device=dev)
        self.d_oseg = torch.empty(shp, dtype=torch.uint8, device=dev)
        self.in_free = None    # compute of the previous use has consumed d_seg / d_seeds
        self.out_done = None   # D2H of the previous use has landed in h_img / h_oseg
        self.params = None


class HostPipeline:
    """``batch`` volumes per slot, ``depth`` slots.  A caller that wants a step of N volumes can run it as
    N / batch micro-steps (``run(steps * N // batch)``): the pipeline then fills and drains in units of
    ``batch`` volumes instead of N, which matters when only a few steps are timed."""

    def __init__(self, generator, batch: int, depth: int = 2, packed_counts=None):
        """``packed_counts``: the inputs are bit-packed seed words (the ``FetalSynthDataset(packed_cache=...)``
        format, sub-class counts ``packed_counts``) instead of four int8 seed volumes per sample; the
        sub-class counts are then drawn per sample and the label volume is unpacked on the device."""
        self.gen = generator
        self.B = batch
        self.shape = tuple(generator.shape)
        self.eng = generator.engine(self.shape)
        dev = self.eng.device
        self.packed_counts = None if packed_counts is None else [int(c) for c in packed_counts]
        wdt = None
        if self.packed_counts is not None:
            from .data.packed import field_layout, word_dtype

            wdt = torch.int16 if np.dtype(word_dtype(field_layout(self.packed_counts))).itemsize == 2 else torch.int32
        self.slots = [_Slot(batch, self.shape, dev, wdt) for _ in range(max(1, depth))]
        self.s_in = torch.cuda.Stream(device=dev)
        self.s_out = torch.cuda.Stream(device=dev)
        self.h2d_bytes = self.slots[0].h_seg.numel() + self.slots[0].h_seeds.numel() * self.slots[0].h_seeds.element_size()
        self.d2h_bytes = self.slots[0].h_img.numel() * 4 + self.slots[0].h_oseg.numel()
        self._next = 0
        self._inflight: deque = deque()

    # ------------------------------------------------------------------ inputs
    def input_buffers(self, slot: int | None = None):
        """Pinned (segmentation [B,*shape] uint8, seeds [B,4,*shape] int8) of the slot the next
        ``submit`` will use: the producer writes its decoded volumes straight into them."""
        s = self.slots[self._next % len(self.slots) if slot is None else slot]
        return s.h_seg, s.h_seeds

    def set_inputs(self, segs, seeds):
        """Copy the same host volumes into every slot (benchmark / tests)."""
        for s in self.slots:
            for b in range(self.B):
                s.h_seg[b].copy_(torch.from_numpy(np.ascontiguousarray(segs[b])))
                for m in range(4):
                    s.h_seeds[b, m].copy_(torch.from_numpy(np.ascontiguousarray(seeds[b][m])))

    def set_inputs_packed(self, segs, words):
        """Copy the same host volumes into every slot: uint8 segmentations and packed seed words (uint16 / uint32)."""
        if self.packed_counts is None:
            raise RuntimeError("HostPipeline was built for int8 seed volumes")
        for s in self.slots:
            for b in range(self.B):
                s.h_seg[b].copy_(torch.from_numpy(np.ascontiguousarray(segs[b])))
                w = np.ascontiguousarray(words[b])
                s.h_seeds[b].copy_(torch.from_numpy(w.view(np.int16 if w.dtype.itemsize == 2 else np.int32)))

    # ------------------------------------------------------------------ pipeline
    def submit(self, scale: bool = True, **kw):
        """Enqueue H2D -> generate -> D2H for the next slot; returns immediately."""
        if len(self._inflight) == len(self.slots):
            raise RuntimeError("HostPipeline: every slot is in flight; call collect() first")
        s = self.slots[self._next % len(self.slots)]
        self._next += 1
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self.s_in):
            if s.in_free is not None:
                self.s_in.wait_event(s.in_free)
            s.d_seg.copy_(s.h_seg, non_blocking=True)
            s.d_seeds.copy_(s.h_seeds, non_blocking=True)
            loaded = torch.cuda.Event()
            loaded.record(self.s_in)
        cur.wait_event(loaded)
        if s.out_done is not None:
            cur.wait_event(s.out_done)  # d_img / d_oseg of this slot are being read by the last D2H
        if self.packed_counts is None:
            seeds = [[s.d_seeds[b, m] for m in range(4)] for b in range(self.B)]
        else:
            from .data.packed import PackedSeeds

            seeds = [PackedSeeds.from_device(s.d_seeds[b], self.packed_counts) for b in range(self.B)]
        _, _, s.params = self.gen.sample_batch([s.d_seg[b] for b in range(self.B)], seeds, scale=scale, out_img=s.d_img, out_seg=s.d_oseg, **kw)
        s.in_free = torch.cuda.Event()
        s.in_free.record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(s.in_free)
            s.h_img.copy_(s.d_img, non_blocking=True)
            s.h_oseg.copy_(s.d_oseg, non_blocking=True)
            s.out_done = torch.cuda.Event()
            s.out_done.record(self.s_out)
        self._inflight.append(s)

    def collect(self):
        """Block until the oldest submitted step is in host memory; returns (image, segmentation,
        params) — views of that slot's pinned buffers, valid until the slot is submitted again."""
        s = self._inflight.popleft()
        s.out_done.synchronize()
        return s.h_img, s.h_oseg, s.params

    def drain(self):
        out = None
        while self._inflight:
            out = self.collect()
        return out

    def step(self, scale: bool = True, **kw):
        """Unpipelined convenience: one submit followed by its collect."""
        self.submit(scale, **kw)
        return self.collect()

    def run(self, steps: int, scale: bool = True, on_result=None, **kw):
        """``steps`` pipelined steps; every result is in host memory when this returns."""
        for _ in range(steps):
            if len(self._inflight) == len(self.slots):
                r = self.collect()
                if on_result is not None:
                    on_result(*r)
            self.submit(scale, **kw)
        while self._inflight:
            r = self.collect()
            if on_result is not None:
                on_result(*r)


class DatasetPipeline:
    """``FetalSynthDataset.sample_batch(indices)`` with host tensors out — the reference's entry point takes a
    subject *index* (``fetalsyngen/data/datasets.py:256``) and returns ``.cpu()`` tensors (``:315-317``).  The
    subject cache (uint8 segmentation + bit-packed seed words) stays resident on the device, so a step moves
    only its indices host->device and its image + segmentation (80 MiB per 256^3 volume) device->host; generation
    of step k+1 overlaps the read-back of step k through ``depth`` rotating buffer sets."""

    def __init__(self, dataset, batch: int, depth: int = 2):
        self.ds, self.B = dataset, batch
        gen = dataset.generator
        self.shape = tuple(gen.shape)
        dev = gen.engine(self.shape).device
        self.slots = [_Slot(batch, self.shape, dev, inputs=False) for _ in range(max(1, depth))]
        self.s_out = torch.cuda.Stream(device=dev)
        self.h2d_bytes = 8 * batch  # the indices
        self.d2h_bytes = self.slots[0].h_img.numel() * 4 + self.slots[0].h_oseg.numel()
        self._next = 0
        self._inflight: deque = deque()

    def submit(self, indices, scale: bool = True, **kw):
        if len(self._inflight) == len(self.slots):
            raise RuntimeError("DatasetPipeline: every slot is in flight; call collect() first")
        if len(indices) != self.B:
            raise ValueError(f"expected {self.B} indices")
        s = self.slots[self._next % len(self.slots)]
        self._next += 1
        cur = torch.cuda.current_stream()
        if s.out_done is not None:
            cur.wait_event(s.out_done)  # d_img / d_oseg of this slot are being read by the last D2H
        out, s.params = self.ds.sample_batch(indices, scale=scale, out_img=s.d_img, out_seg=s.d_oseg, **kw)
        s.names = out["name"]
        done = torch.cuda.Event()
        done.record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            s.h_img.copy_(s.d_img, non_blocking=True)
            s.h_oseg.copy_(s.d_oseg, non_blocking=True)
            s.out_done = torch.cuda.Event()
            s.out_done.record(self.s_out)
        self._inflight.append(s)

    def collect(self):
        """Block until the oldest submitted step is in host memory: (image, segmentation, params) — views of
        that slot's pinned buffers, valid until the slot is submitted again."""
        s = self._inflight.popleft()
        s.out_done.synchronize()
        return s.h_img, s.h_oseg, s.params

    def run(self, index_batches, scale: bool = True, on_result=None, **kw):
        for indices in index_batches:
            if len(self._inflight) == len(self.slots):
                r = self.collect()
                if on_result is not None:
                    on_result(*r)
            self.submit(indices, scale, **kw)
        while self._inflight:
            r = self.collect()
            if on_result is not None:
                on_result(*r)
