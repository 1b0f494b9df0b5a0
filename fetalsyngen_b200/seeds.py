"""Seed generation on the GPU: host mirror of the reference's
``scripts/generate_seeds.py`` over the K6 entry points of libfsg (``csrc/seeds.cu``).

The reference fuses the segmentation labels into four meta-labels (CSF, GM, WM, non-brain tissue),
then, for every requested number of sub-classes n >= 2, clusters the image intensities of each
meta-label with ``sklearn.mixture.GaussianMixture(n, n_init=5, init_params="k-means++")`` and
stores ``10 * meta-label + cluster`` as an int8 volume per meta-label
(``generate_seeds.py:133-211``); one subject with ``--max_subclasses 10`` is 36 fits x 5
initialisations on a pool of host processes.  Here the 180 initialisations of a subject are jobs of
one launch sequence: ordered partition of the voxels by meta-label, k-means++ seeding, EM with the
convergence test on the device, best-of-five selection from one small read-back, and a predict pass
that scatters the labels into the seed volumes.  There is no CPU fallback.

``SeedGenerator.split_labels`` mirrors ``split_lables``; ``process_subject`` / ``main`` mirror the
script's functions and flags, with the file layout the generator's dataset expects
(``subclasses_{n}/{sub}/anat/{label file stem}_mlabel_{m}.nii.gz``).
"""
from __future__ import annotations

import argparse
from pathlib import Path

import numpy as np
import torch

from . import _lib
from .utils.nifti import read_nifti, write_nifti

FETA2META = {1: 1, 4: 1, 2: 2, 6: 2, 5: 3, 7: 3, 3: 3}        # generate_seeds.py:74
DHCP2META = {1: 1, 5: 1, 2: 2, 7: 2, 9: 2, 3: 3, 6: 3, 8: 3}  # generate_seeds.py:84
EM_MAXK = 16
_OTHER = 255  # uint8 code of a label value that is not an integer in 0..254: never matches a table entry


def _stream():
    return torch.cuda.current_stream().cuda_stream


def label_lut(annotation: str = "feta", label2meta: dict | None = None) -> bytes:
    """256-entry table label -> meta-label for ``fsg_seed_partition``; 4 marks the labels that count as
    background (label 0, and the skull label 4 of dHCP, generate_seeds.py:144-146), which become
    meta-label 4 wherever the image is non-zero (:198)."""
    if label2meta is None:
        if annotation not in ("feta", "dhcp"):
            raise ValueError("Unknown annotation type. Should be either 'feta' or 'dhcp'")
        label2meta = FETA2META if annotation == "feta" else DHCP2META
    lut = np.zeros(256, dtype=np.uint8)
    for lab, m in label2meta.items():
        if not (0 < int(lab) < _OTHER and 1 <= int(m) <= 3):
            raise ValueError(f"label {lab} -> meta-label {m} is outside the supported ranges")
        lut[int(lab)] = int(m)
    lut[0] = 4
    if annotation == "dhcp":
        lut[4] = 4
    return lut.tobytes()


def labels_to_u8(segmentation) -> np.ndarray:
    """float / integer label volume -> uint8 codes: NaN -> 0 (generate_seeds.py:141), values that are not
    integers in 0..254 -> 255 (they match no entry of the label table, as in the reference's ``==`` tests)."""
    seg = np.asarray(segmentation.cpu() if isinstance(segmentation, torch.Tensor) else segmentation)
    if seg.dtype == np.uint8:
        return np.ascontiguousarray(seg)
    s = np.nan_to_num(seg.astype(np.float32), nan=0.0, posinf=-1.0, neginf=-1.0)
    ok = (s >= 0) & (s < _OTHER) & (s == np.floor(s))
    return np.ascontiguousarray(np.where(ok, s, _OTHER).astype(np.uint8))


class SeedGenerator:
    """``GaussianMixture`` defaults of the reference call: n_init=5, max_iter=100, tol=1e-3, reg_covar=1e-6."""

    def __init__(self, annotation: str = "feta", device: str = "cuda:0", n_init: int = 5, max_iter: int = 100, tol: float = 1e-3, reg_covar: float = 1e-6,
                 seed: int | None = None, label2meta: dict | None = None):
        _lib.load()
        if not torch.cuda.is_available():
            raise _lib.FsgError("SeedGenerator needs a CUDA device: libfsg has no CPU fallback")
        self.device = torch.device(device)
        self.lut = label_lut(annotation, label2meta)
        self.n_init, self.max_iter, self.tol, self.reg_covar = int(n_init), int(max_iter), float(tol), float(reg_covar)
        # the reference draws from numpy's global, unseeded generator: no stream to reproduce
        self.seed = int(np.random.SeedSequence().generate_state(1, np.uint64)[0]) if seed is None else int(seed)
        self._calls = 0
        self.last_fit: dict = {}

    # ------------------------------------------------------------------ device steps
    def partition(self, image, segmentation):
        """-> (x [total] f32, index [total] i32, counts[4]) with the meta-labels' values back to back."""
        img = torch.as_tensor(np.ascontiguousarray(image, dtype=np.float32) if not isinstance(image, torch.Tensor) else image).to(self.device, torch.float32).contiguous()
        seg = torch.from_numpy(labels_to_u8(segmentation)).to(self.device)
        if tuple(img.shape) != tuple(seg.shape):
            raise ValueError(f"image {tuple(img.shape)} and segmentation {tuple(seg.shape)} differ in shape")
        n = img.numel()
        with torch.cuda.device(self.device):
            x = torch.empty(n, dtype=torch.float32, device=self.device)
            index = torch.empty(n, dtype=torch.int32, device=self.device)
            counts = torch.zeros(4, dtype=torch.int64, device=self.device)
            ws_bytes = _lib.load().fsg_seed_partition_workspace(n)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
            _lib.call("fsg_seed_partition", img.data_ptr(), seg.data_ptr(), self.lut, n, x.data_ptr(), index.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
            c = counts.cpu().tolist()
        return x, index, c

    def fit_jobs(self, specs, seeds_in=None):
        """specs: list of (x tensor, k).  Runs ``n_init`` initialisations of every spec (or one from the
        given sample indices, ``seeds_in[i]``: parity tests) and returns, per spec, the best one:
        dict(weights, means, covariances, n_iter, converged, lower_bound, init, seeds)."""
        ninit = 1 if seeds_in is not None else self.n_init
        njobs = len(specs) * ninit
        dev = self.device
        with torch.cuda.device(dev):
            params = torch.zeros((njobs, 3 * EM_MAXK), dtype=torch.float64, device=dev)
            trace = torch.zeros((njobs, self.max_iter + 1), dtype=torch.float64, device=dev)
            state = torch.zeros((njobs, 2), dtype=torch.int32, device=dev)
            seeds = torch.zeros((njobs, EM_MAXK), dtype=torch.int32, device=dev)
            if seeds_in is not None:
                host = np.zeros((njobs, EM_MAXK), dtype=np.int32)
                for i, s in enumerate(seeds_in):
                    host[i, : len(s)] = np.asarray(s, dtype=np.int32)
                seeds.copy_(torch.from_numpy(host))
            jobs = (_lib.EmJob * njobs)()
            self._calls += 1
            for i, (x, k) in enumerate(specs):
                if not 1 <= k <= EM_MAXK:
                    raise ValueError(f"n_components = {k} outside [1, {EM_MAXK}]")
                if x.numel() < k:
                    raise ValueError(f"Expected n_samples >= n_components but got n_components = {k}, n_samples = {x.numel()}")
                for r in range(ninit):
                    j = jobs[i * ninit + r]
                    q = i * ninit + r
                    j.x, j.n, j.k = x.data_ptr(), x.numel(), k
                    j.params, j.trace, j.state, j.seeds = params[q].data_ptr(), trace[q].data_ptr(), state[q].data_ptr(), seeds[q].data_ptr()
                    j.rng_seed, j.rng_stream = self.seed & 0xFFFFFFFFFFFFFFFF, (self._calls << 32) | q
            lib = _lib.load()
            ws_bytes = lib.fsg_em_workspace(njobs)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            if seeds_in is None:
                _lib.call("fsg_em_seed", jobs, njobs, ws.data_ptr(), ws_bytes, _stream())
            _lib.call("fsg_em_fit", jobs, njobs, self.max_iter, self.tol, self.reg_covar, ws.data_ptr(), ws_bytes, _stream())
            st, tr, pr, sd = state.cpu().numpy(), trace.cpu().numpy(), params.cpu().numpy(), seeds.cpu().numpy()
        out = []
        for i, (x, k) in enumerate(specs):
            best, best_lb = None, -np.inf
            for r in range(ninit):
                q = i * ninit + r
                lb = tr[q, st[q, 0]]
                if best is None or lb > best_lb:  # mixture/_base.py:282: strictly greater, first initialisation on ties
                    best, best_lb = q, lb
            out.append({"weights": pr[best, :k].copy(), "means": pr[best, EM_MAXK : EM_MAXK + k].copy(), "covariances": pr[best, 2 * EM_MAXK : 2 * EM_MAXK + k].copy(),
                        "n_iter": int(st[best, 0]), "converged": bool(st[best, 1]), "lower_bound": float(best_lb), "init": best - i * ninit, "seeds": sd[best, :k].copy(),
                        "trace": tr[best, 1 : st[best, 0] + 1].copy()})
        return out

    def predict_jobs(self, specs, fits, outs=None, want_labels=False):
        """specs: list of (x, k, index, label_base); fits: matching ``fit_jobs`` results (None for k == 1).
        Scatters label_base + component into ``outs[i]`` (int8 volume) and/or returns uint8 labels."""
        njobs = len(specs)
        dev = self.device
        with torch.cuda.device(dev):
            params = np.zeros((njobs, 3 * EM_MAXK), dtype=np.float64)
            for i, f in enumerate(fits):
                if f is not None:
                    k = len(f["weights"])
                    params[i, :k], params[i, EM_MAXK : EM_MAXK + k], params[i, 2 * EM_MAXK : 2 * EM_MAXK + k] = f["weights"], f["means"], f["covariances"]
            params_d = torch.from_numpy(params).to(dev)
            labels = [torch.empty(s[0].numel(), dtype=torch.uint8, device=dev) if want_labels else None for s in specs]
            jobs = (_lib.EmJob * njobs)()
            for i, (x, k, index, base) in enumerate(specs):
                j = jobs[i]
                j.x, j.n, j.k, j.label_base = x.data_ptr(), x.numel(), k, base
                j.index = None if index is None else index.data_ptr()
                j.params = params_d[i].data_ptr()
                j.labels = None if labels[i] is None else labels[i].data_ptr()
                j.out = None if outs is None else outs[i].data_ptr()
            ws_bytes = _lib.load().fsg_em_workspace(njobs)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.call("fsg_em_predict", jobs, njobs, ws.data_ptr(), ws_bytes, _stream())
            torch.cuda.current_stream().synchronize()  # params_d / ws are released when this returns
        return labels

    # ------------------------------------------------------------------ reference surface
    def split_labels(self, image, segmentation, subclasses) -> dict:
        """``split_lables`` (generate_seeds.py:190-211) for one number of sub-classes or a list of them:
        {n_subclasses: {meta-label: int8 device tensor of the image's shape}}."""
        sub_list = [int(subclasses)] if np.isscalar(subclasses) else [int(s) for s in subclasses]
        shape = tuple(segmentation.shape)
        x, index, counts = self.partition(image, segmentation)
        off = np.concatenate([[0], np.cumsum(counts)])
        parts = [(x[off[m] : off[m + 1]], index[off[m] : off[m + 1]]) for m in range(4)]
        fit_specs, where = [], []
        for s in sub_list:
            if s < 1:
                raise ValueError("subclasses must be >= 1")
            for m in range(4):
                if s > 1:
                    if counts[m] < s:
                        raise ValueError(f"Expected n_samples >= n_components but got n_components = {s}, n_samples = {counts[m]} (meta-label {m + 1})")
                    where.append((s, m))
                    fit_specs.append((parts[m][0], s))
        fits = dict(zip(where, self.fit_jobs(fit_specs))) if fit_specs else {}
        self.last_fit = fits
        out = {s: {m + 1: torch.zeros(shape, dtype=torch.int8, device=self.device) for m in range(4)} for s in sub_list}
        pred_specs, pred_fits, outs = [], [], []
        for s in sub_list:
            for m in range(4):
                if counts[m] == 0:
                    continue
                pred_specs.append((parts[m][0], s, parts[m][1], 10 * (m + 1)))
                pred_fits.append(fits.get((s, m)))
                outs.append(out[s][m + 1])
        if pred_specs:
            self.predict_jobs(pred_specs, pred_fits, outs)
        return out

    def process_subject(self, image_path, label_path, out_path, sub_name: str, subclasses, session: str = "") -> list:
        """``process_subject`` (generate_seeds.py:130-172) for all requested sub-class counts at once;
        returns the files written."""
        image = read_nifti(image_path).astype(np.float32)
        label, affine = read_nifti(label_path, with_affine=True)
        res = self.split_labels(image, label, subclasses)
        stem = Path(label_path).name
        stem = stem[: -len(".nii.gz")] if stem.endswith(".nii.gz") else Path(stem).stem
        written = []
        for n_sub, per_label in res.items():
            suffix = f"subclasses_{n_sub}/{sub_name}/anat/" if session == "" else f"subclasses_{n_sub}/{sub_name}/{session}/anat/"
            folder = Path(out_path) / suffix
            folder.mkdir(parents=True, exist_ok=True)
            for m, vol in per_label.items():
                f = folder / f"{stem}_mlabel_{m}.nii.gz"
                write_nifti(f, vol.cpu().numpy(), affine)
                written.append(f)
        return written


    def pack_subject(self, image_path, label_path, out_file, subclasses):
        """Seeds of one subject straight into the bit-packed cache format (``data/packed.py``) that
        ``FetalSynthDataset(packed_cache=...)`` loads — no intermediate NIfTI files."""
        from .data.packed import pack_seed_volumes, save_packed

        image = read_nifti(image_path).astype(np.float32)
        label, affine = read_nifti(label_path, with_affine=True)
        res = self.split_labels(image, label, subclasses)
        words, counts = pack_seed_volumes({n: {m: v.cpu().numpy() for m, v in per.items()} for n, per in res.items()})
        save_packed(out_file, labels_to_u8(label), words, counts, affine)
        return Path(out_file)


def main(argv=None) -> int:
    """Same flags as the reference script (generate_seeds.py:32-59) plus --device / --seed."""
    p = argparse.ArgumentParser(description="Generate seeds for FetalSynthGen (GPU)",
                                epilog="Example: python tools/generate_seeds.py --bids_path /path/to/bids --out_path /path/to/out --max_subclasses 6 --annotation feta")
    p.add_argument("--bids_path", type=str, required=True, help="Path to BIDS folder with the segmentations and images for seeds generation")
    p.add_argument("--out_path", type=str, required=True, help="Path to save the seeds")
    p.add_argument("--max_subclasses", type=int, default=10, help="How many subclasses to simulate for each tissue type (meta-label)")
    p.add_argument("--annotation", type=str, required=True, choices=["feta", "dhcp"], help="Annotation type. Should be either 'feta' or 'dhcp'")
    p.add_argument("--packed", action="store_true", help="write one bit-packed cache file per subject (<out_path>/<sub>.fsgpack.npz) instead of NIfTI seed volumes")
    p.add_argument("--device", type=str, default="cuda:0")
    p.add_argument("--seed", type=int, default=None)
    a = p.parse_args(argv)
    bids = Path(a.bids_path).absolute()
    subjects = sorted(bids.glob("sub-*"))
    print(f"Found {len(subjects)} subjects in {bids}")
    gen = SeedGenerator(a.annotation, a.device, seed=a.seed)
    for sub in subjects:
        imgs = list(sub.glob("**/anat/*_T2w.nii.gz"))[0]
        label = list(sub.glob("**/anat/*_dseg.nii.gz"))[0]
        if a.packed:
            f = gen.pack_subject(imgs, label, Path(a.out_path) / f"{sub.name}.fsgpack.npz", range(1, int(a.max_subclasses) + 1))
            print(f"{sub.name}: {f.name}, {f.stat().st_size / 2**20:.1f} MiB")
            continue
        files = gen.process_subject(imgs, label, a.out_path, sub.name, range(1, int(a.max_subclasses) + 1))
        print(f"{sub.name}: {len(files)} seed volumes")
    return 0
