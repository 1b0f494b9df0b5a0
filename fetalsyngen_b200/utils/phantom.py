"""Procedural label phantom (nested ellipsoids) with the label structure of the reference's
bundled data: a segmentation with labels 0..7 and four seed volumes whose labels are
``10*m + k`` (meta-label m = 1..4, sub-class k), disjoint supports, ~14 % foreground.
Used for synthetic benchmark/test inputs — the reference's NIfTI files do not travel."""
from __future__ import annotations

import numpy as np


def label_phantom(shape, n_sub=(3, 4, 2, 5), seed: int = 0, fill: float = 0.6):
    rs = np.random.RandomState(seed)
    g = np.meshgrid(*[np.linspace(-1, 1, s, dtype=np.float32) for s in shape], indexing="ij", sparse=True)
    r = np.sqrt(sum((gi / (fill + 0.04 * a)) ** 2 for a, gi in enumerate(g)))
    rings = np.array([0.30, 0.50, 0.64, 0.76, 0.86, 0.94, 1.0], dtype=np.float32)
    seg = (7 - np.digitize(r, rings)).clip(0, 7).astype(np.uint8)
    meta = np.array([0, 1, 1, 2, 3, 3, 4, 4], dtype=np.int8)[seg]
    seeds = []
    for m in range(1, 5):
        # smooth-ish sub-class texture: coarse random field, nearest up-sampled
        coarse = rs.randint(0, n_sub[m - 1], size=[max(s // 8, 1) for s in shape]).astype(np.int8)
        idx = np.ix_(*[np.minimum(np.arange(s) // 8, c - 1) for s, c in zip(shape, coarse.shape)])
        sub = coarse[idx]
        seeds.append(np.where(meta == m, 10 * m + sub, 0).astype(np.int8))
    return seg, seeds
