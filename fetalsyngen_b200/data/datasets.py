"""Datasets with the reference's public interface (``fetalsyngen/data/datasets.py:17-370``):
``FetalDataset`` (BIDS discovery), ``FetalTestDataset`` and ``FetalSynthDataset`` whose
``sample / __getitem__ / sample_with_meta`` drive ``FetalSynthGen`` and return
``{"image": (1,H,W,D) float32 in [0,1] on the CPU, "label": (1,H,W,D) int64, "name"}``.

Differences from the reference, all on the host/device plumbing side:
  * decoded segmentations / seeds are cached on the device as uint8 / int8 (the reference
    gunzips four seed files and the segmentation for every sample, ~0.23 s);
  * ``ScaleIntensity`` and the dtype casts run in libfsg kernels;
  * ``sample_batch`` (not in the reference) generates several samples with batched launches and
    leaves them on the device for a co-located trainer.
"""
from __future__ import annotations

import time
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch

from ..generator.model import FetalSynthGen
from ..utils.image_reading import SimpleITKReader
from ..utils.lru import ByteLRU


class BidsIndex:
    """One walk over a BIDS tree: ``entries[(subject, session)]`` lists the ``anat`` files of that scan.

    The reference globs the tree once per (subject, suffix) (``datasets.py:74-95``); here the directory is read
    once and suffix look-ups are dictionary scans, which matters for seed trees (24 look-ups per subject)."""

    def __init__(self, root, subjects=None):
        self.root = Path(root)
        want = None if subjects is None else set(subjects)
        self.entries: dict = {}
        for sub_dir in sorted(d for d in self.root.glob("sub-*") if d.is_dir()):
            if want is not None and sub_dir.name not in want:
                continue
            for child in sorted(c for c in sub_dir.iterdir() if c.is_dir()):
                # <sub>/anat (no session) or <sub>/<ses>/anat
                ses, anat = (None, child) if child.name == "anat" else (child.name, child / "anat")
                self.entries[(sub_dir.name, ses)] = sorted(anat.glob("*.nii.gz")) if anat.is_dir() else []

    def keys(self):
        return sorted(self.entries, key=lambda k: (k[0], k[1] or ""))

    def find(self, key, suffix):
        """Files of scan ``key`` named ``<sub>[_<ses>]*_<suffix>.nii.gz``."""
        sub, ses = key
        stem = sub if ses is None else f"{sub}_{ses}"
        tail = f"_{suffix}.nii.gz"
        return [f for f in self.entries.get(key, []) if f.name.startswith(stem) and f.name.endswith(tail)]


class FetalDataset:
    """Subject discovery for the datasets below.  Public attributes follow the reference's ``FetalDataset``
    (``datasets.py:17-113``): ``subjects``, ``sub_ses``, ``img_paths``, ``segm_paths``, ``loader``."""

    def __init__(self, bids_path: str, sub_list: list[str] | None):
        self.bids_path = Path(bids_path)
        self._index = BidsIndex(self.bids_path, sub_list)
        self.sub_ses = self._index.keys()
        self.subjects = sorted({sub for sub, _ in self.sub_ses})
        self.loader = SimpleITKReader()
        self.img_paths = self._one_per_scan(self._index, "T2w", required=getattr(self, "_needs_images", True))
        self.segm_paths = self._one_per_scan(self._index, "dseg")

    @staticmethod
    def _sub_ses_string(sub, ses):
        return sub if ses is None else f"{sub}_{ses}"

    def _sub_ses_idx(self, idx):
        return self._sub_ses_string(*self.sub_ses[idx])

    def _one_per_scan(self, index: BidsIndex, suffix: str, required: bool = True):
        """Exactly one file per scan with the given suffix, in ``sub_ses`` order; ``required=False`` maps a
        missing file to ``None`` (images are optional when intensities are synthesised from seeds).  Errors as
        in the reference: ``FileNotFoundError`` for none, ``RuntimeError`` for several (``datasets.py:85-94``)."""
        out = []
        for key in self.sub_ses:
            hits = index.find(key, suffix)
            if len(hits) > 1:
                raise RuntimeError(f"Multiple files found for requested subject {key[0]} in {index.root}: {[f.name for f in hits]}")
            if not hits and required:
                raise FileNotFoundError(f"No files found for requested subject {key[0]} in {index.root} (suffix {suffix})")
            out.append(hits[0] if hits else None)
        return out

    def __len__(self):
        return len(self.subjects)

    def __getitem__(self, idx):
        raise NotImplementedError("This method should be implemented in the child class.")


class FetalTestDataset(FetalDataset):
    """Real images + segmentations for validation / testing (``datasets.py:116-198``): returns
    ``{"image": (1,H,W,D), "label": (1,H,W,D) int64, "name"}`` with optional dictionary transforms."""

    def __init__(self, bids_path: str, sub_list: list[str] | None, transforms=None):
        super().__init__(bids_path, sub_list)
        self.transforms = transforms

    def __getitem__(self, idx) -> dict:
        image, segm = self.loader(self.img_paths[idx]), self.loader(self.segm_paths[idx])
        if image.dim() not in (3, 4):
            raise ValueError(f"Expected 3D or 4D image, got {image.dim()}D image.")
        if image.dim() == 3:
            image, segm = image[None], segm[None]
        data = {"image": image, "label": segm.long(), "name": self._sub_ses_idx(idx)}
        if self.transforms:
            data = self.transforms(data)
            data["label"] = data["label"].long()
        return data

    def reverse_transform(self, data: dict) -> dict:
        return self.transforms.inverse(data) if self.transforms else data


class FetalSynthDataset(FetalDataset):
    """On-the-fly generation / augmentation of fetal images (datasets.py:201-370)."""

    def __init__(self, bids_path: str, generator: FetalSynthGen, seed_path: str | None, sub_list: list[str] | None,
                 load_image: bool = False, image_as_intensity: bool = False, packed_cache: str | None = None):
        """``packed_cache``: directory of bit-packed subject files (``data/packed.py``).  A missing file is
        written from the subject's NIfTIs on first use (the one-time conversion); afterwards a process
        start reads one file per subject instead of decoding 25 gzip volumes."""
        self._needs_images = bool(load_image or image_as_intensity)
        super().__init__(bids_path, sub_list)
        self.seed_path = Path(seed_path) if isinstance(seed_path, (str, Path)) else None
        self.load_image = load_image
        self.generator = generator
        self.image_as_intensity = image_as_intensity
        self.packed_cache = Path(packed_cache) if packed_cache is not None else None
        self._init_caches()
        if not self.image_as_intensity and isinstance(self.seed_path, Path):
            if not self.seed_path.exists():
                raise FileNotFoundError(f"Provided seed path {self.seed_path} does not exist.")
            self._load_seed_path()

    @classmethod
    def from_packed(cls, packed_cache: str, generator: FetalSynthGen, sub_list: list[str] | None = None) -> "FetalSynthDataset":
        """Dataset over a directory of bit-packed subject files alone (``<subject>[_<session>].fsgpack.npz``, written by
        ``tools/pack_dataset.py`` or on first use of ``packed_cache=``): no BIDS tree and no NIfTI decoding at all —
        what a training job ships to its nodes.  ``sample`` / ``sample_batch`` / ``DeviceBatchLoader`` work as usual."""
        self = cls.__new__(cls)
        self.bids_path, self.seed_path = None, Path(packed_cache)
        self.packed_cache = Path(packed_cache)
        files = sorted(self.packed_cache.glob("*.fsgpack.npz"))
        names = [f.name[: -len(".fsgpack.npz")] for f in files]
        if sub_list is not None:
            names = [n for n in names if n.split("_ses-")[0] in set(sub_list)]
        if not names:
            raise FileNotFoundError(f"No *.fsgpack.npz subject files under {packed_cache}")
        self.sub_ses = [tuple(n.split("_", 1)) if "_ses-" in n else (n, None) for n in names]
        self.subjects = sorted({sub for sub, _ in self.sub_ses})
        self.loader = SimpleITKReader()
        self.img_paths = [None] * len(self.sub_ses)
        self.segm_paths = [None] * len(self.sub_ses)
        self.seed_paths = {}
        self.load_image = self.image_as_intensity = self._needs_images = False
        self.generator = generator
        self._init_caches()
        return self

    def _init_caches(self, budget_gb: float | None = None):
        """Device-resident subject caches (uint8 segmentation: 16 MiB, packed seed words: 32 MiB per 256^3 subject),
        least recently used subjects dropped beyond ``FSG_SUBJECT_CACHE_GB`` (default 32) per process."""
        import os

        gb = float(os.environ.get("FSG_SUBJECT_CACHE_GB", "32")) if budget_gb is None else budget_gb
        self._seg_cache = ByteLRU(int(gb * 2**30 / 3))
        self._packed = ByteLRU(int(gb * 2**30 * 2 / 3))

    def _load_seed_path(self):
        """{sub_ses: {n_subclasses: {meta_label: path}}} (datasets.py:246-270)."""
        self.seed_paths = {self._sub_ses_string(sub, ses): defaultdict(dict) for (sub, ses) in self.sub_ses}
        avail = [int(x.name.replace("subclasses_", "")) for x in self.seed_path.glob("subclasses_*")]
        if not avail:
            raise FileNotFoundError(f"No subclasses_* folders under {self.seed_path}")
        wanted = sorted({sub for sub, _ in self.sub_ses})
        for n_sub in range(min(avail), max(avail) + 1):
            seed_path = self.seed_path / f"subclasses_{n_sub}"
            if not seed_path.exists():
                raise FileNotFoundError(f"Provided seed path {seed_path} does not exist.")
            index = BidsIndex(seed_path, wanted)
            for i in range(1, 5):
                files = self._one_per_scan(index, f"mlabel_{i}")
                for (sub, ses), file in zip(self.sub_ses, files):
                    self.seed_paths[self._sub_ses_string(sub, ses)][n_sub][i] = file

    # ------------------------------------------------------------------ device caches
    def _packed_subject(self, idx):
        """(uint8 segmentation, PackedSeeds) of subject idx from the packed cache, converting on first use."""
        hit = self._packed.get(idx)
        if hit is None:
            from .packed import load_packed, pack_subject

            name = self._sub_ses_string(*self.sub_ses[idx])
            f = self.packed_cache / f"{name}.fsgpack.npz"
            if not f.exists():
                if self.segm_paths[idx] is None:
                    raise FileNotFoundError(f"{f} is missing and there is no BIDS tree to convert it from")
                pack_subject(self.segm_paths[idx], self.seed_paths[name], f)
            seg, seeds, _ = load_packed(f)
            hit = (seg, seeds)
            self._packed.put(idx, hit, seg.nbytes + int(np.prod(seeds.shape)) * seeds.word_bytes)
        return hit

    def _seeds(self, idx):
        """What the generator receives as ``seeds``: the reference's {n_subclasses: {meta_label: path}} or the packed cache."""
        if self.packed_cache is not None:
            return self._packed_subject(idx)[1]
        return self.seed_paths[self._sub_ses_string(*self.sub_ses[idx])]

    def _segmentation(self, idx) -> torch.Tensor:
        """uint8 label map of subject idx, decoded once and kept on the generator's device."""
        seg = self._seg_cache.get(idx)
        if seg is None and self.packed_cache is not None and not self.image_as_intensity and self.seed_path is not None:
            raw = torch.from_numpy(self._packed_subject(idx)[0])
            seg = raw.to(self.generator.engine(tuple(raw.shape)).device).contiguous()
            self._seg_cache.put(idx, seg, seg.numel())
        if seg is None:
            raw = torch.nan_to_num(self.loader(self.segm_paths[idx]).float())
            if raw.numel() and (float(raw.min()) < 0 or float(raw.max()) > 255 or not torch.equal(raw, raw.round())):
                raise ValueError(f"{self.segm_paths[idx]}: labels are not integers in 0..255 (uint8 label maps only)")
            seg = raw.to(torch.uint8).to(self.generator.engine(tuple(raw.shape)).device).contiguous()
            self._seg_cache.put(idx, seg, seg.numel())
        return seg

    # ------------------------------------------------------------------ reference API
    def sample(self, idx, genparams: dict = {}) -> tuple[dict, dict]:
        generation_params = {}
        image = self.loader(self.img_paths[idx]) if self.load_image else None
        segm = self._segmentation(idx)
        name = self._sub_ses_string(*self.sub_ses[idx])
        seeds = None
        if self.seed_path is not None:
            seeds = self._seeds(idx)
        if self.image_as_intensity:
            seeds = None
        generation_params["idx"] = idx
        generation_params["img_paths"] = str(self.img_paths[idx])
        generation_params["segm_paths"] = str(self.img_paths[idx])
        generation_params["seeds"] = str(self.seed_path)
        t0 = time.time()
        gen_output, segmentation, image, synth_params = self.generator.sample(image=image, segmentation=segm, seeds=seeds, genparams=genparams)
        eng = self.generator.engine(tuple(gen_output.shape))
        gen_output = eng.scale_intensity(gen_output.contiguous())
        image = eng.scale_intensity(image.contiguous()) if image is not None else None
        label = eng.from_u8(segmentation if segmentation.dtype == torch.uint8 else eng.to_u8(segmentation), torch.int64)
        gen_output, label, image = self._to_host(gen_output, label, image)
        generation_params = {**generation_params, **synth_params}
        generation_params["generation_time"] = time.time() - t0
        return {"image": gen_output.unsqueeze(0), "label": label.unsqueeze(0), "name": name}, generation_params

    @staticmethod
    def _to_host(*tensors):
        """Device -> host copies of the results (datasets.py:315-317 `.cpu()`).  The host tensors are
        fresh allocations from torch's caching *pinned* allocator: a D2H copy into pageable memory
        bounces through a staging buffer at ~10 GB/s and first-touches 192 MB of new pages per sample
        (measured ~100 ms), the pinned copy runs at the link rate (~3.5 ms)."""
        out = []
        for t in tensors:
            if t is None:
                out.append(None)
                continue
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            out.append(h)
        torch.cuda.current_stream().synchronize()
        return out

    def __getitem__(self, idx) -> dict:
        data_out, generation_params = self.sample(idx)
        self.generation_params = generation_params
        return data_out

    def sample_with_meta(self, idx: int, genparams: dict = {}) -> dict:
        data, generation_params = self.sample(idx, genparams=genparams)
        data["generation_params"] = generation_params
        return data

    # ------------------------------------------------------------------ device fast path
    def sample_batch(self, indices, scale: bool = True, out_img=None, out_seg=None, sample_ids=None, base_seed: int | None = None, artifacts: bool = False):
        """Generate ``len(indices)`` samples with batched launches.  Returns
        ``{"image": (B,1,H,W,D) float32, "label": (B,1,H,W,D) uint8, "name": [...]}`` on the
        generator's device plus the list of per-sample parameter dictionaries.  ``artifacts=True``: the generator's
        SR artifacts are applied to every sample like ``__getitem__`` does (ScaleIntensity after them).

        All parameters of the batch are drawn at once (``batch_draw.py``) as a function of (base seed, sample id).
        Without ``sample_ids`` the dataset numbers its samples itself and takes the base seed from numpy's global
        generator on first use, so ``np.random.seed(s)`` before the first call still fixes the whole stream."""
        if self.image_as_intensity or self.seed_path is None:
            raise ValueError("sample_batch needs seed-based intensity generation")
        if sample_ids is None:
            if getattr(self, "_batch_seed", None) is None:
                self._batch_seed, self._batch_next = int(np.random.randint(0, 2**31 - 1)), 0
            sample_ids = list(range(self._batch_next, self._batch_next + len(indices)))
            self._batch_next += len(indices)
            base_seed = self._batch_seed
        segs = [self._segmentation(i) for i in indices]
        names = [self._sub_ses_string(*self.sub_ses[i]) for i in indices]
        img, seg, params = self.generator.sample_batch(segs, [self._seeds(i) for i in indices], scale=scale, out_img=out_img, out_seg=out_seg, sample_ids=sample_ids,
                                                       base_seed=int(base_seed or 0), artifacts=artifacts)
        return {"image": img.unsqueeze(1), "label": seg.unsqueeze(1), "name": names}, params


class DeviceBatchLoader:
    """GPU-resident replacement of ``DataLoader(FetalSynthDataset, num_workers=k)`` for an on-the-fly
    training loop on the same device (the reference hands CPU tensors from spawned workers to the
    trainer, ``fetalsyngen/test_dl.py:17-23``, because "GPU memory cannot be easily shared between
    processes", ``docs/datasets.md:4-6``).

    Iterating yields ``{"image": (B,1,H,W,D) float32 in [0,1], "label": (B,1,H,W,D) uint8 or int64,
    "name": [...], "params": [...]}`` on the generator's device.  Batches are produced on a private
    CUDA stream one step ahead of the consumer into ``depth`` rotating buffer sets; the consumer's
    stream waits on the producing event only, nothing is copied to the host.  A yielded batch stays
    valid until ``depth - 1`` further batches have been requested.
    """

    def __init__(self, dataset: "FetalSynthDataset", batch_size: int, num_batches: int | None = None, shuffle: bool = True, labels_int64: bool = False,
                 depth: int = 2, base_seed: int | None = None, rank: int = 0, world: int = 1):
        if dataset.image_as_intensity or dataset.seed_path is None:
            raise ValueError("DeviceBatchLoader needs seed-based intensity generation")
        if depth < 2:
            raise ValueError("depth must be >= 2 (one batch in use, one in flight)")
        self.ds, self.B, self.depth = dataset, int(batch_size), int(depth)
        self.num_batches = num_batches if num_batches is not None else max(1, len(dataset) // self.B)
        self.shuffle, self.labels_int64 = shuffle, labels_int64
        self.base_seed, self.rank, self.world = base_seed, rank, world
        self.epoch = 0  # advanced by every ``__iter__`` (or ``set_epoch``): epochs draw different samples
        gen = dataset.generator
        self.shape = tuple(gen.shape)
        self.eng = gen.engine(self.shape)
        dev = self.eng.device
        self.stream = torch.cuda.Stream(device=dev)
        shp = (self.B, *self.shape)
        self._img = [torch.empty(shp, dtype=torch.float32, device=dev) for _ in range(self.depth)]
        self._seg = [torch.empty(shp, dtype=torch.uint8, device=dev) for _ in range(self.depth)]
        self._lab = [torch.empty(shp, dtype=torch.int64, device=dev) for _ in range(self.depth)] if labels_int64 else None
        self._released = [None] * self.depth  # event: the consumer is done with this slot

    def __len__(self):
        return self.num_batches

    def _indices(self, step, rs):
        n = len(self.ds)
        return [int(v) for v in (rs.randint(0, n, self.B) if self.shuffle else [(step * self.B + k) % n for k in range(self.B)])]

    def set_epoch(self, epoch: int):
        """Reproducibility contract with ``base_seed``: sample k of step s of epoch e has id
        ``(e * num_batches + s) * B * world + k * world + rank`` and every draw is a function of (base_seed, id), so a
        run is reproducible, independent of the sharding, and no two epochs repeat a sample.  The subject indices of
        an epoch come from ``RandomState((base_seed, rank, epoch))``."""
        self.epoch = int(epoch)

    def _produce(self, step, slot, rs, epoch):
        from ..sharding import step_ids

        idx = self._indices(step, rs)
        names = [self.ds._sub_ses_string(*self.ds.sub_ses[i]) for i in idx]
        kw = {}
        if self.base_seed is not None:  # reproducible, sharding-independent sample streams
            kw = {"sample_ids": step_ids(epoch * self.num_batches + step, self.B, self.rank, self.world), "base_seed": self.base_seed}
        with torch.cuda.stream(self.stream):
            # first-use uploads of a subject (segmentation, packed words) are enqueued on the producer stream too
            segs = [self.ds._segmentation(i) for i in idx]
            if self._released[slot] is not None:
                self.stream.wait_event(self._released[slot])
            _, _, params = self.ds.generator.sample_batch(segs, [self.ds._seeds(i) for i in idx], scale=True, out_img=self._img[slot], out_seg=self._seg[slot], **kw)
            if self._lab is not None:
                self.eng._call("fsg_u8_to_i64", self._seg[slot].data_ptr(), self._lab[slot].data_ptr(), self._seg[slot].numel())
            done = torch.cuda.Event()
            done.record(self.stream)
        return done, names, params

    def __iter__(self):
        epoch = self.epoch
        self.epoch += 1
        rs = np.random.RandomState(None if self.base_seed is None else [self.base_seed & 0xFFFFFFFF, self.rank, epoch])
        pending = self._produce(0, 0, rs, epoch)
        for step in range(self.num_batches):
            slot = step % self.depth
            done, names, params = pending
            if step + 1 < self.num_batches:
                pending = self._produce(step + 1, (step + 1) % self.depth, rs, epoch)
            cur = torch.cuda.current_stream()
            cur.wait_event(done)
            label = self._lab[slot] if self._lab is not None else self._seg[slot]
            yield {"image": self._img[slot].unsqueeze(1), "label": label.unsqueeze(1), "name": names, "params": params}
            rel = torch.cuda.Event()
            rel.record(cur)  # everything the consumer queued on its stream so far has read this slot
            self._released[slot] = rel
