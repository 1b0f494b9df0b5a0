#!/usr/bin/env python
"""BASELINE.json configs[4]: the batched generator feeding an on-the-fly 3-D U-Net training loop on
the same GPU (generator throughput vs training-step consumption).

The U-Net is a plain-PyTorch consumer written for this measurement (it is not a component of the
reference); the reference's counterpart is ``DataLoader(FetalSynthDataset)`` with CPU tensors
(``fetalsyngen/test_dl.py:11-30``).  Here generated batches never leave the device: the generator
runs on its own CUDA stream one batch ahead of the optimiser step.

    python tools/train_unet.py --steps 20 --crop 128 --train-batch 4
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from fetalsyngen_b200.sharding import step_ids  # noqa: E402
from fetalsyngen_b200.utils.phantom import label_phantom  # noqa: E402


def block(cin, cout):
    return nn.Sequential(nn.Conv3d(cin, cout, 3, padding=1, bias=False), nn.GroupNorm(8, cout), nn.LeakyReLU(0.01, inplace=True),
                         nn.Conv3d(cout, cout, 3, padding=1, bias=False), nn.GroupNorm(8, cout), nn.LeakyReLU(0.01, inplace=True))


class UNet3D(nn.Module):
    def __init__(self, cin=1, ncls=8, ch=(16, 32, 64, 128)):
        super().__init__()
        self.enc = nn.ModuleList([block(cin if i == 0 else ch[i - 1], c) for i, c in enumerate(ch)])
        self.up = nn.ModuleList([nn.ConvTranspose3d(ch[i], ch[i - 1], 2, stride=2) for i in range(len(ch) - 1, 0, -1)])
        self.dec = nn.ModuleList([block(2 * ch[i - 1], ch[i - 1]) for i in range(len(ch) - 1, 0, -1)])
        self.head = nn.Conv3d(ch[0], ncls, 1)

    def forward(self, x):
        skips = []
        for i, e in enumerate(self.enc):
            x = e(x)
            if i < len(self.enc) - 1:
                skips.append(x)
                x = F.max_pool3d(x, 2)
        for up, dec in zip(self.up, self.dec):
            x = dec(torch.cat([up(x), skips.pop()], 1))
        return self.head(x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, default=256)
    ap.add_argument("--crop", type=int, default=128)
    ap.add_argument("--gen-batch", type=int, default=8)
    ap.add_argument("--train-batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    dev = "cuda:0"
    shape = (args.shape,) * 3
    seg_h, seeds_h = label_phantom(shape)
    gen = bench.build_generator(shape, dev)
    seg_d = torch.from_numpy(seg_h).to(dev)
    seeds_d = [torch.from_numpy(s).to(dev) for s in seeds_h]
    B, TB, C = args.gen_batch, args.train_batch, args.crop
    net = UNet3D().to(dev).to(memory_format=torch.channels_last_3d)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    gstream = torch.cuda.Stream()
    bufs = [(torch.empty((B, *shape), dtype=torch.float32, device=dev), torch.empty((B, *shape), dtype=torch.uint8, device=dev)) for _ in range(2)]
    counter = [0]

    def generate(slot):
        ids = step_ids(counter[0], B, 0, 1)
        counter[0] += 1
        gen.sample_batch([seg_d] * B, [seeds_d] * B, scale=True, out_img=bufs[slot][0], out_seg=bufs[slot][1], sample_ids=ids, base_seed=1234)

    def crops(slot, k):
        """TB random crops out of the B generated volumes of a slot."""
        img, seg = bufs[slot]
        xs, ys = [], []
        for t in range(TB):
            b = (k * TB + t) % B
            o = np.random.randint(0, args.shape - C + 1, 3)
            xs.append(img[b, o[0] : o[0] + C, o[1] : o[1] + C, o[2] : o[2] + C])
            ys.append(seg[b, o[0] : o[0] + C, o[1] : o[1] + C, o[2] : o[2] + C])
        x = torch.stack(xs).unsqueeze(1).contiguous(memory_format=torch.channels_last_3d)
        return x, torch.stack(ys).long()

    def train_step(x, y):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.cross_entropy(net(x), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    def timeit(fn, n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            fn(i)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n

    # generator alone / trainer alone
    for i in range(args.warmup):
        generate(0)
    t_gen = timeit(lambda i: generate(0), args.steps)
    x0, y0 = crops(0, 0)
    for i in range(args.warmup):
        train_step(x0, y0)
    t_train = timeit(lambda i: train_step(x0, y0), args.steps)

    # on-the-fly: one generated batch feeds B // TB optimiser steps; the next batch is generated on
    # a side stream while those steps run
    per = max(1, B // TB)
    losses = []
    generate(0)
    with torch.cuda.stream(gstream):  # engines are per stream: build the producer stream's engine (scratch, tables) before the clock starts
        gstream.wait_stream(torch.cuda.current_stream())
        generate(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nsteps = 0
    for it in range(args.steps):
        slot = it % 2
        with torch.cuda.stream(gstream):
            gstream.wait_stream(torch.cuda.current_stream())  # the slot being overwritten was consumed two iterations ago
            generate(1 - slot)
        for k in range(per):
            x, y = crops(slot, k)
            losses.append(train_step(x, y))
            nsteps += 1
        torch.cuda.current_stream().wait_stream(gstream)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    first, last = float(torch.stack(losses[:per]).mean()), float(torch.stack(losses[-per:]).mean())
    print(json.dumps({
        "config": "configs[4]: generator feeding an on-the-fly 3-D U-Net loop on one GPU", "shape": args.shape, "crop": C, "gen_batch": B, "train_batch": TB,
        "generator_alone_volumes_per_s": B / t_gen, "trainer_alone_steps_per_s": 1 / t_train, "trainer_alone_volumes_per_s": TB / t_train,
        "on_the_fly_steps_per_s": nsteps / dt, "on_the_fly_generated_volumes_per_s": args.steps * B / dt,
        "slowdown_vs_trainer_alone": (nsteps / dt) / (1 / t_train), "loss_first": first, "loss_last": last,
        "unet_params": sum(p.numel() for p in net.parameters()), "precision": "bf16 autocast, channels_last_3d",
    }))


if __name__ == "__main__":
    main()
