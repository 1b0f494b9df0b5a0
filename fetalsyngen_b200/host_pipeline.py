"""Host-buffer front end of the batched generator: what a data-loading process calls when the
inputs (segmentation + seed volumes) and the outputs (image + segmentation) live in host
memory, as in the reference's ``FetalSynthDataset.sample`` (``fetalsyngen/data/datasets.py:
256-327``: host tensors in, ``.cpu()`` tensors out).  Pinned staging buffers, one copy stream
per direction, events instead of host synchronisation between the stages."""
from __future__ import annotations

import numpy as np
import torch


class HostPipeline:
    def __init__(self, generator, batch: int):
        self.gen = generator
        self.B = batch
        self.shape = tuple(generator.shape)
        self.eng = generator.engine(self.shape)
        dev = self.eng.device
        shp = (batch, *self.shape)
        self.h_seg = torch.empty(shp, dtype=torch.uint8, pin_memory=True)
        self.h_seeds = torch.empty((batch, 4, *self.shape), dtype=torch.int8, pin_memory=True)
        self.h_img = torch.empty(shp, dtype=torch.float32, pin_memory=True)
        self.h_oseg = torch.empty(shp, dtype=torch.uint8, pin_memory=True)
        self.d_seg = torch.empty(shp, dtype=torch.uint8, device=dev)
        self.d_seeds = torch.empty((batch, 4, *self.shape), dtype=torch.int8, device=dev)
        self.d_img = torch.empty(shp, dtype=torch.float32, device=dev)
        self.d_oseg = torch.empty(shp, dtype=torch.uint8, device=dev)
        self.s_in = torch.cuda.Stream(device=dev)
        self.s_out = torch.cuda.Stream(device=dev)
        self.h2d_bytes = self.h_seg.numel() + self.h_seeds.numel()
        self.d2h_bytes = self.h_img.numel() * 4 + self.h_oseg.numel()

    def set_inputs(self, segs, seeds):
        for b in range(self.B):
            self.h_seg[b].copy_(torch.from_numpy(np.ascontiguousarray(segs[b])))
            for m in range(4):
                self.h_seeds[b, m].copy_(torch.from_numpy(np.ascontiguousarray(seeds[b][m])))

    def step(self, scale: bool = True):
        """H2D inputs -> generate -> D2H outputs; returns after the outputs are in host memory."""
        cur = torch.cuda.current_stream()
        self.s_in.wait_stream(cur)
        with torch.cuda.stream(self.s_in):
            self.d_seg.copy_(self.h_seg, non_blocking=True)
            self.d_seeds.copy_(self.h_seeds, non_blocking=True)
        cur.wait_stream(self.s_in)
        _, _, params = self.gen.sample_batch([self.d_seg[b] for b in range(self.B)], [[self.d_seeds[b, m] for m in range(4)] for b in range(self.B)], scale=scale, out_img=self.d_img, out_seg=self.d_oseg)
        self.s_out.wait_stream(cur)
        with torch.cuda.stream(self.s_out):
            self.h_img.copy_(self.d_img, non_blocking=True)
            self.h_oseg.copy_(self.d_oseg, non_blocking=True)
        self.s_out.synchronize()
        return self.h_img, self.h_oseg, params
