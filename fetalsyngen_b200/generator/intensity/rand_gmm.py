"""Seed meta-labels -> GMM intensity image.  Same class name, constructor and method
signatures as the reference's ``ImageFromSeeds``
(``fetalsyngen/generator/intensity/rand_gmm.py:9-154``); the per-voxel work runs in the
``fsg_gmm`` kernel (seed sum + label->(mu,sigma) lookup + noise + clamp in one pass)."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Iterable

import numpy as np
import torch

from ...utils.lru import ByteLRU
from ...utils.nifti import read_nifti, to_ras


class ImageFromSeeds:
    def __init__(
        self,
        min_subclusters: int,
        max_subclusters: int,
        seed_labels: Iterable[int],
        generation_classes: Iterable[int],
        meta_labels: int = 4,
    ):
        self.min_subclusters = min_subclusters
        self.max_subclusters = max_subclusters
        seed_labels, generation_classes = list(seed_labels), list(generation_classes)
        if len(set(seed_labels)) != len(seed_labels):
            raise ValueError("Parameter seed_labels should have unique values.")
        if len(seed_labels) != len(generation_classes):
            raise ValueError("Parameters seed_labels and generation_classes should have the same lengths.")
        self.seed_labels = seed_labels
        self.generation_classes = generation_classes
        self.meta_labels = meta_labels
        # (path, device) -> int8 device volume; least recently used volumes are dropped beyond the byte budget
        # (24 volumes = 384 MiB per 256^3 subject; the bit-packed subject cache needs 32 MiB instead)
        self._cache = ByteLRU(int(float(os.environ.get("FSG_SEED_CACHE_GB", "16")) * 2**30))

    # ------------------------------------------------------------------ host draws
    def draw_subclusters(self, mlabel2subclusters=None, genparams: dict = {}) -> dict:
        """{meta_label: n_subclusters}, drawn like rand_gmm.py:81-87."""
        if mlabel2subclusters is None:
            mlabel2subclusters = {m: np.random.randint(self.min_subclusters, self.max_subclusters + 1) for m in range(1, self.meta_labels + 1)}
        if "mlabel2subclusters" in genparams.keys():
            mlabel2subclusters = genparams["mlabel2subclusters"]
        return mlabel2subclusters

    def draw_gmm(self, genparams: dict = {}, inject: dict | None = None):
        """mus / sigmas tables (float32 numpy) in the reference's draw order (rand_gmm.py:116-145)."""
        inject = inject or {}
        nlabels = max(self.seed_labels) + 1
        nsamp = len(self.seed_labels)

        def u(key, n):
            return np.asarray(inject[key], dtype=np.float32) if key in inject else torch.rand(n, dtype=torch.float).numpy()

        if "mus" in genparams.keys():
            mus = torch.as_tensor(genparams["mus"]).detach().cpu().numpy().astype(np.float32).copy()
        else:
            mus = (np.float32(25) + np.float32(200) * u("mus_u", nlabels)).astype(np.float32)
        if "sigmas" in genparams.keys():
            sigmas = torch.as_tensor(genparams["sigmas"]).detach().cpu().numpy().astype(np.float32).copy()
        else:
            sigmas = (np.float32(5) + np.float32(20) * u("sigmas_u", nlabels)).astype(np.float32)
        if self.generation_classes != self.seed_labels:
            pert = np.asarray(inject["mus_perturb"], dtype=np.float32) if "mus_perturb" in inject else torch.randn(nsamp, dtype=torch.float).numpy()
            tied = mus[np.asarray(self.generation_classes)] + np.float32(25) * pert
            mus[np.asarray(self.seed_labels)] = np.clip(tied, np.float32(0), np.float32(225))
        return mus, sigmas

    # ------------------------------------------------------------------ seed cache
    def seed_volume(self, path, device) -> torch.Tensor:
        """Decoded int8 seed volume resident on ``device`` (decoded once per path)."""
        key = (str(path), str(device))
        vol = self._cache.get(key)
        if vol is None:
            arr, _ = to_ras(*read_nifti(path, with_affine=True))  # rand_gmm.py:91-96: seeds are reoriented to RAS too
            if arr.dtype != np.int8:
                if arr.min() < -128 or arr.max() > 127:
                    raise ValueError(f"{path}: seed labels do not fit int8")
                arr = arr.astype(np.int8)
            vol = torch.from_numpy(np.ascontiguousarray(arr)).to(device)
            self._cache.put(key, vol, vol.numel() * vol.element_size())
        return vol

    def select_seeds(self, seeds, mlabel2subclusters: dict, device) -> list[torch.Tensor]:
        """The volumes whose sum is the sample's label map: the four selected seed files, or — for a
        bit-packed subject cache (``data/packed.py``) — the one label volume unpacked on the device."""
        from ...data.packed import PackedSeeds

        if isinstance(seeds, PackedSeeds):
            return [seeds.labels(mlabel2subclusters, device)]
        return [self.seed_volume(seeds[mlabel2subclusters[m]][m], device) for m in range(1, self.meta_labels + 1)]

    # ------------------------------------------------------------------ reference-compatible API
    def load_seeds(self, seeds: dict, mlabel2subclusters: dict | None = None, genparams: dict = {}, device=None):
        """Summed seed label map as a LongTensor (rand_gmm.py:51-99)."""
        m2s = self.draw_subclusters(mlabel2subclusters, genparams)
        vols = self.select_seeds(seeds, m2s, device if device is not None else "cuda")
        total = vols[0].to(torch.int64)
        for v in vols[1:]:
            total = total + v
        return total, {"mlabel2subclusters": m2s}

    def sample_intensities(self, seeds: torch.Tensor, device: str, genparams: dict = {}, inject: dict | None = None):
        """Label map -> intensities (rand_gmm.py:101-154)."""
        from ...engine import SamplePlan, engine_for

        eng = engine_for(device, tuple(seeds.shape))
        mus, sigmas = self.draw_gmm(genparams, inject)
        lab = seeds.to(eng.device)
        if lab.dtype not in (torch.int8, torch.uint8):
            lab = lab.to(torch.uint8)
        lab = lab.contiguous()
        plan = SamplePlan(mus=mus, sigmas=sigmas, rng_seed=int(torch.randint(0, 2**62, (1,)).item()))
        if inject and "gmm_noise" in inject:
            plan.gmm_noise = torch.as_tensor(inject["gmm_noise"], dtype=torch.float32).to(eng.device).contiguous()
        out = torch.empty((1, lab.numel()), dtype=torch.float32, device=eng.device)
        eng.gmm([plan], [[lab.view(-1)]], out)
        return out.view(lab.shape), {"mus": torch.from_numpy(mus).to(eng.device), "sigmas": torch.from_numpy(sigmas).to(eng.device)}

