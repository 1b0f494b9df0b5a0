#!/usr/bin/env python
"""Benchmark of the FetalSynthGen per-sample generation path (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels)
    python bench.py --impl reference --steps K --warmup W    # CPU implementation of the reference path

Inputs: the reference's bundled 256^3 subjects sub-sta21 / sub-sta30 / sub-sta38 (BASELINE.json configs[0..1]),
committed bit-packed under tests/golden/subjects (uint8 segmentation + one uint16 word per voxel holding the 24 seed
volumes).  Sample id k uses subject k mod 3 and draws its own sub-class counts, so every volume of a step has its own
label map.

One *step* = one batch of `--batch` (default 8) volumes per GPU through the whole base pipeline (GMM -> warp + gamma +
bias -> blur + down-sample + noise -> up-sample /max -> ScaleIntensity), all stage probabilities forced to 1 (fixed
work), Philox noise.
  value              volumes/s, subject cache resident in HBM (device-timed, max over ranks)
  e2e                the same through the host-buffer API (`HostPipeline`): every step copies its segmentations + int8
                     seed volumes from pinned host memory and its images + segmentations back (80 MiB each way per volume)
  e2e_packed_inputs  host-buffer leg with the seeds in the bit-packed word format (48 MiB in per volume)
  e2e_dataset_cache  `FetalSynthDataset.sample_batch(indices)` through `DatasetPipeline`: the reference's entry point
                     takes a subject index (datasets.py:256); the subject cache stays on the device, indices go in,
                     image + segmentation come out to pinned host memory
Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "256^3 synth volumes/sec"
DOMINANT = "fsg_warp"  # largest share of the step (profiles/*_launch_shares.txt)
UNIT = "volumes/s"
SUBJECTS = ("sub-sta21", "sub-sta30", "sub-sta38")
SUBJECT_DIR = ROOT / "tests" / "golden" / "subjects"
MIN_TIMED_S = 0.5  # the timed region is stretched to at least this long (more steps than --steps when needed)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU while the timed region runs.  NVML polled every 2 ms from a
    thread (handle looked up by UUID, then PCI bus id, then index); when NVML is unusable or yields no sample,
    `nvidia-smi -lms 20` on the same GPU.  Errors are kept in the record instead of being swallowed."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.nvml, self._stop = index, [], None, None, False
        self.sm, self.mask, self.max_mhz, self.errors, self.smi_id = [], 0, None, [], str(index)

    def _nvml_handle(self, pynvml):
        import torch

        props = torch.cuda.get_device_properties(self.index)
        uuid = getattr(props, "uuid", None)
        if uuid is not None:
            try:
                self.smi_id = f"GPU-{uuid}"
                return pynvml.nvmlDeviceGetHandleByUUID(f"GPU-{uuid}".encode())
            except Exception as e:
                self.errors.append(f"nvml by uuid: {type(e).__name__}: {e}")
        bus = getattr(props, "pci_bus_id", None)
        if bus is not None:
            for i in range(pynvml.nvmlDeviceGetCount()):
                h = pynvml.nvmlDeviceGetHandleByIndex(i)
                if int(pynvml.nvmlDeviceGetPciInfo(h).bus) == int(bus):
                    return h
        visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = self.index
        if visible and all(v.strip().isdigit() for v in visible.split(",")):
            phys = int(visible.split(",")[self.index])
        return pynvml.nvmlDeviceGetHandleByIndex(phys)

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = self._nvml_handle(pynvml)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))  # fails here, not silently in the thread
            self.nvml = (pynvml, h)
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception as e:
            self.errors.append(f"nvml: {type(e).__name__}: {e}")
            self.nvml = None
        self._start_smi()

    def _start_smi(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", self.smi_id], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception as e:
            self.errors.append(f"nvidia-smi: {type(e).__name__}: {e}")
            self.proc = None

    def _poll(self):
        pynvml, h = self.nvml
        fails = 0
        while not self._stop:
            try:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception as e:
                fails += 1
                if fails == 1:
                    self.errors.append(f"nvml poll: {type(e).__name__}: {e}")
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def _smi_summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return sm, mx, reasons

    def stop(self) -> dict:
        out = None
        if self.nvml is not None:
            self._stop = True
            time.sleep(0.005)
            if self.sm:
                reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
                out = {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.sm), "source": "nvml, 2 ms poll"}
        if out is None and self.proc is not None:
            time.sleep(0.1)
            self.proc.terminate()
            sm, mx, reasons = self._smi_summary()
            out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else self.max_mhz, "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20"}
        if out is None or not out.get("samples"):
            # last resort: one synchronous query (the GPU is still warm right after the timed region)
            try:
                r = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", self.smi_id], capture_output=True, text=True, timeout=20)
                self.rows = [[c.strip() for c in line.split(",")] for line in r.stdout.splitlines()]
                sm, mx, reasons = self._smi_summary()
                out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else self.max_mhz, "reasons": sorted(reasons), "samples": len(sm),
                       "source": "nvidia-smi, one query after the timed region"}
            except Exception as e:
                self.errors.append(f"nvidia-smi query: {type(e).__name__}: {e}")
                out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["clock sampling unavailable"], "samples": 0}
        if self.errors:
            out["errors"] = self.errors[:4]
        return out


def bind_to_gpu_numa_node(local: int) -> str:
    """Pin this rank to the CPUs of its GPU's NUMA node before any pinned host memory is allocated:
    the pipeline's staging buffers then live next to the GPU's PCIe root."""
    try:
        import pynvml

        pynvml.nvmlInit()
        import torch

        props = torch.cuda.get_device_properties(local)
        h = None
        for i in range(pynvml.nvmlDeviceGetCount()):
            hi = pynvml.nvmlDeviceGetHandleByIndex(i)
            if hasattr(props, "pci_bus_id") and int(pynvml.nvmlDeviceGetPciInfo(hi).bus) == int(props.pci_bus_id):
                h = hi
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        node = int(Path(f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node").read_text())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"numa node {node}, {len(cpus)} cpus"
        return f"numa node {node}: no allowed cpu there"
    except Exception as e:  # affinity is an optimisation only
        return f"numa: not bound ({type(e).__name__})"


# ---------------------------------------------------------------------------------- inputs
def load_subjects(shape):
    """[(name, uint8 segmentation, packed seed words, counts)] of the bundled subjects; other extents than the
    fixtures' 256^3 are nearest-index scalings of them."""
    out = []
    for name in SUBJECTS:
        with np.load(SUBJECT_DIR / f"{name}.fsgpack.npz") as z:
            seg, words, counts = z["seg"], z["words"], z["counts"].tolist()
        if tuple(seg.shape) != tuple(shape):
            idx = [np.floor((np.arange(s) + 0.5) * seg.shape[a] / s).astype(int) for a, s in enumerate(shape)]
            seg, words = np.ascontiguousarray(seg[np.ix_(*idx)]), np.ascontiguousarray(words[np.ix_(*idx)])
        out.append((name, seg, words, counts))
    return out


def reference_motion_extension():
    """Baseline leg of the motion comparison: the reference's own slice-acquisition extension, built from the
    reference's sources into oracle/_ref by oracle/build_ref.py (None when it is absent).  It is only timed next to
    libfsg, never called by the product."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import build_ref

    return build_ref.load_built()


def draw_oracle_params(rs, shape, res=0.5):
    """Drawn parameters for one sample in np_oracle.generate_base form (all gates on)."""
    from fetalsyngen_b200.tables import make_affine_matrix, resample_size, resample_stds

    q = {"mus": (25 + 200 * rs.rand(50)).astype(np.float32), "sigmas": (5 + 20 * rs.rand(50)).astype(np.float32), "gmm_noise": rs.randn(*shape).astype(np.float32), "flip": bool(rs.rand() < 0.5), "resolution": np.array([res] * 3)}
    rot = (2 * 20 * rs.rand(3) - 20) / 180 * np.pi
    q["A"] = make_affine_matrix(rot, 0.04 * rs.rand(3) - 0.02, 1 + 0.2 * rs.rand(3) - 0.1).astype(np.float32)
    q["c2"] = (np.array(shape) - 1) / 2
    s = [int(round((0.03 + 0.03 * rs.rand()) * v)) for v in shape]
    q["Fsmall"] = (4 * rs.rand() * rs.randn(*s, 3)).astype(np.float32)
    q["gamma"] = float(np.exp(0.1 * rs.randn()))
    b = [max(int(round((0.004 + 0.016 * rs.rand()) * v)), 1) for v in shape]
    q["bf_low"] = ((0.01 + 0.29 * rs.rand()) * rs.randn(*b)).astype(np.float32)
    sp = res + 2 * res * rs.rand()
    q["spacing"] = np.array([sp] * 3)
    q["stds"] = resample_stds(q["spacing"], [res] * 3, rs.rand())
    q["noise_std"] = float(5 + 10 * rs.rand())
    q["noise"] = rs.randn(*[resample_size(v, res, sp) for v in shape]).astype(np.float32)
    return q


def _cpu_worker(args):
    shape, seed, nvol = args
    sys.path.insert(0, str(ROOT / "oracle"))
    import np_oracle as O
    from fetalsyngen_b200.data.packed import unpack_numpy

    name, seg, words, counts = load_subjects(shape)[seed % len(SUBJECTS)]
    rs = np.random.RandomState(seed)
    t = 0.0
    for _ in range(nvol):
        lab = unpack_numpy(words, counts, {m: int(rs.randint(1, 7)) for m in range(1, 5)})
        q = draw_oracle_params(rs, shape)
        t0 = time.perf_counter()
        out, sg, _ = O.generate_base(lab, seg, q)
        O.scale_intensity(out)
        t += time.perf_counter() - t0
    return t


def cpu_port_throughput(shape, workers: int, vols_per_worker: int, pool=None):
    """volumes/s of the numpy port of the reference path on `workers` host processes."""
    t0 = time.perf_counter()
    if workers == 1 or pool is None:
        _cpu_worker((shape, 0, vols_per_worker))
        workers = 1
    else:
        pool.map(_cpu_worker, [(shape, i, vols_per_worker) for i in range(workers)])
    dt = time.perf_counter() - t0
    return workers * vols_per_worker / dt, dt


def host_workers():
    cores = os.cpu_count() or 1
    try:
        mem_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2**30
    except Exception:
        mem_gb = 16
    return max(1, min(cores, 32, int(mem_gb // 6)))


# ---------------------------------------------------------------------------------- reference arm
def reference_proper_root():
    """Tree holding the reference's own generator, if one is reachable from this host: /root/reference (build
    container) or an installed copy under baseline/_ref.  `pip install --target baseline/_ref /root/reference`
    succeeds but installs two files only — the reference's setup.py lists packages=["fetalsyngen"] without its
    sub-packages — so on the GPU box (which only receives /root/repo) this returns None and the pinned port is timed."""
    for root in (os.environ.get("FSG_REFERENCE_ROOT"), "/root/reference", str(ROOT / "baseline" / "_ref")):
        if root and (Path(root) / "fetalsyngen" / "generator" / "model.py").exists() and (Path(root) / "configs" / "dataset" / "generator" / "default.yaml").exists():
            return Path(root)
    return None


def reference_proper_throughput(root: Path, shape, steps: int):
    """volumes/s of the UNMODIFIED reference (torch, device='cpu', all host threads) on sub-sta30 written back to
    NIfTI seed files, all stage probabilities 1.  Uses the import stubs of oracle/ref_import.py."""
    import tempfile

    import torch

    os.environ["FSG_REFERENCE_ROOT"] = str(root)
    sys.path.insert(0, str(ROOT / "oracle"))
    import ref_import
    from fetalsyngen_b200.data.packed import unpack_numpy
    from fetalsyngen_b200.utils.nifti import write_nifti

    ref_import.load_reference()
    cfg = ref_import.reference_generator_config(device="cpu", shape=shape)
    gen = ref_import.instantiate(cfg)
    for obj, attr in ((gen.spatial_deform, "prob"), (gen.resampled, "prob"), (gen.biasfield, "prob"), (gen.gamma, "prob"), (gen.noise, "prob")):
        setattr(obj, attr, 1.0)
    name, seg, words, counts = load_subjects(shape)[1]
    tmp = Path(tempfile.mkdtemp(prefix="fsg_ref_"))
    seeds = {}
    for n in counts:
        lab = unpack_numpy(words, counts, {m: n for m in range(1, 5)})
        seeds[n] = {}
        for m in range(1, 5):
            v = np.where(lab // 10 == m, lab, 0).astype(np.int8)
            seeds[n][m] = tmp / f"s{n}_m{m}.nii.gz"
            write_nifti(seeds[n][m], v)
    torch.set_num_threads(os.cpu_count() or 1)
    np.random.seed(1234)
    torch.manual_seed(1234)
    segt = torch.from_numpy(seg.astype(np.float32))
    mt = sys.modules["monai.transforms"]
    gen.sample(image=None, segmentation=segt, seeds=seeds)  # warm-up (file cache, thread pool)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = gen.sample(image=None, segmentation=segt, seeds=seeds)[0]
        mt.ScaleIntensity(0, 1)(out)
    dt = time.perf_counter() - t0
    return steps / dt, dt, torch.get_num_threads()


def run_reference(args, shape):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    root = reference_proper_root()
    if root is not None:
        try:
            value, dt, threads = reference_proper_throughput(root, shape, args.steps)
            line = {
                "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": 1,
                "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"base pipeline, all stage probs=1, {shape[0]}^3 @0.5mm sub-sta30; one step = 1 volume through the unmodified FetalSynthGen.sample + ScaleIntensity on the CPU", "shape": list(shape)},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": f"{args.steps} volumes, the reference's own torch implementation from {root} (device='cpu', {threads} torch threads; the four seed NIfTI reads of every sample included, as in the reference's generation_time)"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
            }
            print(json.dumps(line))
            return
        except Exception as e:  # fall through to the pinned port
            print(f"bench.py: reference proper at {root} could not run ({type(e).__name__}: {e}); timing the port", file=sys.stderr)
    import multiprocessing as mp

    workers = host_workers()
    pool = mp.get_context("spawn").Pool(workers) if workers > 1 else None
    cpu_port_throughput((32, 32, 32), workers, 1, pool)  # worker start-up, imports, page-in (untimed)
    vals = []
    t_all = 0.0
    for _ in range(args.steps):
        v, dt = cpu_port_throughput(shape, workers, 1, pool)
        vals.append(v)
        t_all += dt
    if pool is not None:
        pool.close()
    value = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000 * t_all / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"base pipeline, all stage probs=1, {shape[0]}^3 @0.5mm sub-sta21/30/38; one step = {workers} volumes (1 per host process)", "shape": list(shape)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": f"{args.steps} x {workers} volumes, numpy port of the reference path (oracle/np_oracle.py, pinned to the unmodified reference at 256^3 by tests/golden/full_*.npz), one process per volume; host has {os.cpu_count()} logical CPUs (workers capped by min(cpus, 32, RAM / 6 GB)); the reference tree itself is not on this host"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------- our arm
def default_artifacts(prob=1.0):
    """The four SR-artifact stages with the reference's default parameters
    (configs/dataset/generator/default.yaml:58-142), probabilities forced to `prob`."""
    from fetalsyngen_b200.generator.artifacts.utils import ReconMergeParams, ReconParams, ScannerParams, StructNoiseMergeParams
    from fetalsyngen_b200.generator.augmentation.artifacts import BlurCortex, SimulatedBoundaries, SimulateMotion, StructNoise

    smp = StructNoiseMergeParams(merge_type="perlin", gauss_nloc_min=5, gauss_nloc_max=15, gauss_sigma_mu=25, gauss_sigma_std=5, perlin_res_list=[1, 2],
                                 perlin_octaves_list=[1, 2, 4], perlin_persistence=0.5, perlin_lacunarity=2, perlin_increase_size=0.1)
    sp = ScannerParams(resolution_slice_fac_min=0.5, resolution_slice_fac_max=2, resolution_slice_max=1.5, slice_thickness_min=1.5, slice_thickness_max=3.5, gap_min=1.5,
                       gap_max=5.5, min_num_stack=2, max_num_stack=6, max_num_slices=250, noise_sigma_min=0, noise_sigma_max=0.1, TR_min=1, TR_max=2, prob_void=0.2,
                       prob_gamma=0.1, gamma_std=0.05, slice_size=None, restrict_transform=False, txy=3.0)
    rmp = ReconMergeParams(merge_type="perlin", perlin_res_list=[1, 2], perlin_octaves_list=[1, 2, 4], perlin_persistence=0.5, perlin_lacunarity=2, gauss_ngaussians_min=2,
                           gauss_ngaussians_max=4, perlin_increase_size=0.25)
    rp = ReconParams(prob_misreg_slice=0.1, slices_misreg_ratio=0.1, prob_misreg_stack=0.1, txy=3.0, prob_merge=1.0, merge_params=rmp, prob_smooth=0.2, prob_rm_slices=0.3,
                     rm_slices_min=0.1, rm_slices_max=0.4)
    return dict(
        blur_cortex=BlurCortex(prob, 2, 50, 200, 3, 1, 2, 1),
        struct_noise=StructNoise(prob, 3, 0.2, 0.4, smp, 1, 5),
        simulate_motion=SimulateMotion(prob, sp, rp),
        boundaries=SimulatedBoundaries(1.0 - prob, 1.0 if prob >= 1 else 0.5, 1.0 if prob >= 1 else 0.5),
    )


def build_generator(shape, device, artifacts=None):
    from fetalsyngen_b200.generator.augmentation.synthseg import RandBiasField, RandGamma, RandNoise, RandResample
    from fetalsyngen_b200.generator.deformation.affine_nonrigid import SpatialDeformation
    from fetalsyngen_b200.generator.intensity.rand_gmm import ImageFromSeeds
    from fetalsyngen_b200.generator.model import FetalSynthGen

    labels = [0] + list(range(10, 50))
    classes = [0] + [10] * 10 + [20] * 10 + [30] * 10 + list(range(40, 50))
    return FetalSynthGen(
        shape=list(shape), resolution=[0.5, 0.5, 0.5], device=device,
        intensity_generator=ImageFromSeeds(1, 6, labels, classes),
        spatial_deform=SpatialDeformation(20, 0.02, 0.1, list(shape), 1.0, True, 0.03, 0.06, 4, 0.5, device),
        resampler=RandResample(1.0, 0.5, 1.5), bias_field=RandBiasField(1.0, 0.004, 0.02, 0.01, 0.3),
        noise=RandNoise(1.0, 5, 15), gamma=RandGamma(1.0, 0.1), **(artifacts or {}),
    )


# algorithmic HBM bytes per output voxel N; f3 = coarse voxels / N of the timed samples
ALGO_BYTES_PER_VOXEL = {
    "fsg_gmm": lambda f3: 2 + 4,               # 2 B of packed seed words read (the subject-cache path) + 4 B written
    "fsg_warp": lambda f3: 10,                 # img 4 + seg 1 read, img 4 + seg 1 written
    "fsg_sepconv": lambda f3: 4 + 4 * f3,      # read N, write n^3 (the three passes move 4N(1 + 2f + 2f^2 + f^3))
    "fsg_zoom_minmax": lambda f3: 4 * f3,      # read n^3
    "fsg_zoom": lambda f3: 4 * f3 + 4,         # read n^3, write N
}


def motion_vs_reference_extension(dev, reps=3):
    """ms of the slice acquisition / PSF reconstruction at the 215-tap shape of profiles/*_motion.jsonl (0.6 mm
    in-plane, 2.5 mm thick, 288^2 slices), libfsg against the reference's own extension (oracle/_ref)."""
    import torch

    from fetalsyngen_b200.generator.artifacts import simulate_reco as SR
    from fetalsyngen_b200.generator.artifacts import svort

    ext = reference_motion_extension()
    S = 256
    seg = load_subjects((S, S, S))[1][1]
    rs = np.random.RandomState(0)
    vol = torch.from_numpy((seg > 0).astype(np.float32) * (0.3 + 0.7 * rs.rand(S, S, S).astype(np.float32))).to(dev)
    res_s, thick, gap = 0.6, 2.5, 3.5
    np.random.seed(1)
    psf = svort.get_PSF(res_ratio=(res_s / 0.5, res_s / 0.5, thick / 0.5))
    ss = int(np.ceil(int(np.sqrt(3 * S * S / 2.0) * 0.5 / res_s) / 32.0) * 32)
    ns = int(S * 0.5 / gap) + 2
    init = svort.random_init_stack_transforms(ns, gap, False, 3.0)
    motion = svort.sample_motion(np.arange(ns) * 1.5, True)
    mat = svort.mat_update_resolution(motion.compose(init).matrix(), 0.5, 0.5)
    mats = np.concatenate([mat] * max(1, min(6, 250 // ns)))[:250]

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    quads = SR.volume_xyquads(vol)
    sl = SR.slice_acquisition(mats, vol, psf, (ss, ss), res_s / 0.5, pairs=quads)
    taps = int((psf != 0).sum())
    out = {"psf_taps": taps, "slice_size": ss, "slices_forward": int(mat.shape[0]), "slices_adjoint": int(mats.shape[0]),
           "forward_ms_ours": timed(lambda: SR.slice_acquisition(mat, vol, psf, (ss, ss), res_s / 0.5, pairs=quads)),
           "adjoint_ms_ours": timed(lambda: SR.slice_acquisition_adjoint(mats, psf, sl, (S, S, S), res_s / 0.5))}
    out["forward_tap_evals_per_s"] = mat.shape[0] * ss * ss * taps / (out["forward_ms_ours"] / 1000)
    out["adjoint_tap_evals_per_s"] = mats.shape[0] * ss * ss * taps / (out["adjoint_ms_ours"] / 1000)
    if ext is not None:
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        empty = torch.empty(0, device=dev)
        tm, tp, tms = t(mat), t(psf), t(mats)
        out["forward_ms_reference_ext"] = timed(lambda: ext.forward(tm, vol[None, None], empty, empty, tp, [ss, ss], float(res_s / 0.5), False, False))
        out["adjoint_ms_reference_ext"] = timed(lambda: ext.adjoint_forward(tms, tp, sl, empty, empty, [S, S, S], float(res_s / 0.5), True, True))
        out["forward_speedup"] = out["forward_ms_reference_ext"] / out["forward_ms_ours"]
        out["adjoint_speedup"] = out["adjoint_ms_reference_ext"] / out["adjoint_ms_ours"]
    else:
        out["reference_ext"] = "oracle/_ref/slice_acq_cuda.so not present"
    return out


def artifacts_throughput(shape, dev, subjects_dev, nsamples=12, batch=2):
    """configs[2]: volumes/s of the full pipeline with the four SR artifacts forced on, through the batched public
    call `FetalSynthGen.sample_batch(..., artifacts=True)` (batched base path, then the artifacts per sample on the
    device, ScaleIntensity last; one stream)."""
    import torch

    gen = build_generator(shape, dev, default_artifacts(1.0))

    def one(k):
        ids = list(range(k * batch, (k + 1) * batch))
        sub = [subjects_dev[i % len(subjects_dev)] for i in ids]
        return gen.sample_batch([s[0] for s in sub], [s[1] for s in sub], scale=True, sample_ids=ids, base_seed=99, artifacts=True)[0]

    nwarm = 3
    for k in range(nwarm):  # lazy module loads, allocator growth, per-shape tables, every artifact branch once
        one(k)
    torch.cuda.synchronize()
    per = []
    for k in range(nwarm, nsamples // batch + nwarm):
        t0 = time.perf_counter()
        one(k)
        torch.cuda.synchronize()
        per.append((time.perf_counter() - t0) / batch)
    dt = sum(per) * batch
    n = len(per) * batch
    return {"value": n / dt, "unit": UNIT, "samples": n, "batch": batch, "ms_per_sample": 1000 * dt / n,
            "ms_per_sample_median": 1000 * float(np.median(per)), "ms_per_sample_max": 1000 * max(per), "ms_each": [round(1000 * t, 1) for t in per],
            "workload": "configs[2]: base pipeline + BlurCortex + StructNoise + SimulateMotion + SimulatedBoundaries, all forced on, sample_batch(artifacts=True), one stream"}


def run_ours(args, shape):
    import torch
    import torch.distributed as dist

    from fetalsyngen_b200 import _lib
    from fetalsyngen_b200.data.packed import PackedSeeds, save_packed, unpack_numpy
    from fetalsyngen_b200.sharding import step_ids

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    B = args.batch
    nvox = int(np.prod(shape))
    np.random.seed(1234 + rank)
    torch.manual_seed(1234 + rank)

    subjects = load_subjects(shape)
    gen = build_generator(shape, dev)
    gen.engine(shape)
    subjects_dev = [(torch.from_numpy(seg).to(dev), PackedSeeds(words, counts, device=dev)) for _, seg, words, counts in subjects]
    out_img = torch.empty((B, *shape), dtype=torch.float32, device=dev)
    out_seg = torch.empty((B, *shape), dtype=torch.uint8, device=dev)
    counter = [0]
    last_params = [None]

    def step():
        # rank r owns sample ids r, r+R, ...; every draw is a function of (1234, sample id), not of R;
        # sample id k reads subject k mod 3 and draws its own sub-class counts (its own label map)
        ids = step_ids(counter[0], B, rank, world)
        counter[0] += 1
        segs = [subjects_dev[i % len(subjects_dev)][0] for i in ids]
        seeds = [subjects_dev[i % len(subjects_dev)][1] for i in ids]
        last_params[0] = gen.sample_batch(segs, seeds, scale=True, out_img=out_img, out_seg=out_seg, sample_ids=ids, base_seed=1234)[2]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step()
    barrier()
    # ---- length of the timed region: at least --steps and at least MIN_TIMED_S (estimated from 3 more steps)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    est = (time.perf_counter() - t0) / 10
    steps_timed = max(args.steps, int(math.ceil(1.15 * MIN_TIMED_S / max(est, 1e-6))))
    if world > 1:
        t = torch.tensor([steps_timed], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        steps_timed = int(t.item())
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.stats.reset()
    # CUDA events bracket the dominant kernel's launches inside the timed region (two event records per step);
    # bracketing every entry point costs ~0.4 ms of host time per step, so the full per-kernel table comes from
    # extra steps after the timed region
    _lib.stats.timing = {DOMINANT}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps_timed):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    _lib.stats.timing = False
    dominant = _lib.stats.elapsed_ms()
    launches = _lib.stats.total_kernels()  # kernels of libfsg launched inside the timed region
    clocks = sampler.stop() if rank == 0 else {}
    # ---- per-entry-point table (10 further steps, every call bracketed) and the coarse-grid fraction of those samples
    _lib.stats.reset()
    _lib.stats.timing = True
    f3s = []
    for _ in range(10):
        step()
        for pr in last_params[0]:
            sp = pr["resample_params"]["spacing"]
            f3s.append(float(np.prod([int(shape[a] * 0.5 / sp[a]) for a in range(3)])) / nvox if sp is not None else 0.0)
    _lib.stats.timing = False
    per_call = _lib.stats.elapsed_ms()
    f3 = float(np.mean(f3s))
    # ---- host cost of a step: wall time to ISSUE steps, in rounds of 4 from an idle device.  The engine's parameter
    # ring lets the host run at most 8 steps ahead of the device, so a longer burst would measure the device.
    host_s = 0.0
    for _ in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            step()
        host_s += time.perf_counter() - t0
    host_ms = 1000 * host_s / 32
    torch.cuda.synchronize()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * steps_timed / (ms / 1000)

    def timed_host_leg(run):
        barrier()
        t0 = time.perf_counter()
        run()
        barrier()
        t = torch.tensor([max(time.perf_counter() - t0, 1e-9)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- e2e through the host-buffer API
    from fetalsyngen_b200.host_pipeline import DatasetPipeline, HostPipeline

    e2e_steps = 0 if args.no_e2e else max(args.steps, 10)
    # a step's B volumes travel as `micro` micro-batches through the pipeline (same bytes per step; the pipeline
    # fills / drains in units of B / micro volumes, so a short timed region is not dominated by the first H2D and
    # the last D2H)
    micro = args.micro if B % args.micro == 0 else 1
    mb = B // micro
    sink = [0.0]

    def consume(h_img, h_seg, params):
        sink[0] += float(h_img[0, 0, 0, 0]) + float(h_seg[-1, -1, -1, -1])  # the host reads the step's result

    e2e_value, e2e_bytes, e2e_packed, e2e_cache, host_dma = 0.0, (0, 0), None, None, None
    ids_of = lambda k, n: [rank + world * (k * n + j) for j in range(n)]
    if e2e_steps:
        hp = HostPipeline(gen, mb, depth=args.depth + 1)
        # every slot holds different volumes: subject (slot + b) mod 3 with its own sub-class counts
        rs = np.random.RandomState(5)
        for si, s in enumerate(hp.slots):
            for b in range(mb):
                _, seg, words, counts = subjects[(si + b) % len(subjects)]
                lab = unpack_numpy(words, counts, {m: int(rs.randint(1, 7)) for m in range(1, 5)})
                s.h_seg[b].copy_(torch.from_numpy(seg))
                for m in range(4):
                    s.h_seeds[b, m].copy_(torch.from_numpy(np.where(lab // 10 == m + 1, lab, 0).astype(np.int8)))
        k = [0]

        def run_host(n):
            for _ in range(n):
                if len(hp._inflight) == len(hp.slots):
                    consume(*hp.collect())
                hp.submit(True, sample_ids=ids_of(k[0], mb), base_seed=1234)
                k[0] += 1
            while hp._inflight:
                consume(*hp.collect())

        run_host((args.depth + 1) * micro)
        e2e_value = world * B * e2e_steps / timed_host_leg(lambda: run_host(e2e_steps * micro))
        e2e_bytes = (hp.h2d_bytes * micro, hp.d2h_bytes * micro)
        del hp
        torch.cuda.empty_cache()

        # ---- same leg with the inputs in the bit-packed subject-cache format: uint8 segmentation + one uint16 word
        # per voxel for all six sub-class counts; the counts are drawn per sample, labels decoded inside fsg_gmm
        if not args.no_packed_e2e:
            hp = HostPipeline(gen, mb, depth=args.depth + 1, packed_counts=subjects[0][3])
            for si, s in enumerate(hp.slots):
                for b in range(mb):
                    _, seg, words, _ = subjects[(si + b) % len(subjects)]
                    s.h_seg[b].copy_(torch.from_numpy(seg))
                    s.h_seeds[b].copy_(torch.from_numpy(words.view(np.int16)))
            run_host((args.depth + 1) * micro)
            tsec = timed_host_leg(lambda: run_host(e2e_steps * micro))
            e2e_packed = {"value": world * B * e2e_steps / tsec, "unit": UNIT, "h2d_bytes_per_step": hp.h2d_bytes * micro, "d2h_bytes_per_step": hp.d2h_bytes * micro,
                          "inputs": "uint8 segmentation + uint16 bit-packed seed words per voxel (subject-cache format), sub-class counts drawn per sample"}
            del hp
            torch.cuda.empty_cache()

        # ---- the reference's real entry point: a subject INDEX in, host tensors out; subject cache on the device
        if not args.no_cache_e2e:
            import tempfile

            from fetalsyngen_b200.data.datasets import FetalSynthDataset

            cache_dir = Path(tempfile.mkdtemp(prefix=f"fsg_cache_r{rank}_"))
            for name, seg, words, counts in subjects:
                save_packed(cache_dir / f"{name}.fsgpack.npz", seg, words, counts)
            ds = FetalSynthDataset.from_packed(cache_dir, gen)
            dp = DatasetPipeline(ds, mb, depth=args.depth + 1)
            kk = [0]

            def run_cache(n):
                batches = []
                for _ in range(n):
                    ids = ids_of(kk[0], mb)
                    kk[0] += 1
                    batches.append(([i % len(ds.sub_ses) for i in ids], ids))
                for idx, ids in batches:
                    if len(dp._inflight) == len(dp.slots):
                        consume(*dp.collect())
                    dp.submit(idx, True, sample_ids=ids, base_seed=1234)
                while dp._inflight:
                    consume(*dp.collect())

            run_cache((args.depth + 1) * micro)
            tsec = timed_host_leg(lambda: run_cache(e2e_steps * micro))
            e2e_cache = {"value": world * B * e2e_steps / tsec, "unit": UNIT, "h2d_bytes_per_step": dp.h2d_bytes * micro, "d2h_bytes_per_step": dp.d2h_bytes * micro,
                         "api": "FetalSynthDataset.from_packed(...).sample_batch(indices) through DatasetPipeline: subject cache resident on the device, image + segmentation to pinned host memory"}
            # aggregate device->host ceiling of the box: D2H copies alone, all ranks at once
            hbuf, dbuf = dp.slots[0].h_img, dp.slots[0].d_img
            hbuf.copy_(dbuf, non_blocking=True)
            reps = 6
            tsec = timed_host_leg(lambda: [hbuf.copy_(dbuf, non_blocking=True) for _ in range(reps)] and torch.cuda.synchronize())
            gbs = world * reps * hbuf.numel() * 4 / tsec / 1e9
            host_dma = {"d2h_GBps_all_ranks": gbs, "ranks": world, "e2e_dataset_cache_ceiling_volumes_per_s": gbs * 1e9 / (nvox * 5), "numa": numa,
                        "note": "pinned-memory D2H copies alone, every rank at once: what the box's host side allows for an output of 5 bytes per voxel"}
            del dp, ds
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: dominant kernel (CUDA-event time inside the timed region) + per-entry-point table
    peak, peak_src = peaks()
    name, (ncalls, tms) = DOMINANT, dominant[DOMINANT]
    algo = ALGO_BYTES_PER_VOXEL[name](f3) * nvox * B
    achieved = algo / (tms / ncalls / 1000) / 1e9 if tms > 0 else 0.0
    traffic_all = {}
    prof = ROOT / "profiles" / "dominant_kernel_traffic.json"
    if prof.exists():
        traffic_all = json.loads(prof.read_text())
    per_kernel = []
    for k_, (n_, t_) in sorted(per_call.items(), key=lambda kv: -kv[1][1]):
        ms_ = t_ / n_
        row = {"name": k_, "ms": round(ms_, 4)}
        if k_ in ALGO_BYTES_PER_VOXEL:
            gb = ALGO_BYTES_PER_VOXEL[k_](f3) * nvox * B / 1e9
            row.update({"algorithmic_GB": round(gb, 4), "GBps": round(gb / (ms_ / 1000), 1), "frac": round(gb / (ms_ / 1000) / peak, 4), "dram_bytes": traffic_all.get(k_)})
        per_kernel.append(row)
    if per_kernel and per_kernel[0]["name"] != DOMINANT:
        print(f"bench.py: note: {DOMINANT} is no longer the slowest entry point: {per_kernel[0]}", file=sys.stderr)

    # ---- CPU baseline: bounded sample of the same workload on the host cores
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, dt = cpu_port_throughput(shape, 1, 1)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"1 volume of the same workload (sub-sta21), numpy port of the reference path (oracle/np_oracle.py, pinned to the unmodified reference at 256^3), {dt:.1f} s, single process"}

    extras = {}
    if world == 1 and not args.no_extras:
        try:
            extras["artifacts_volumes_per_s"] = artifacts_throughput(shape, dev, subjects_dev)
        except Exception as e:
            extras["artifacts_volumes_per_s"] = {"error": f"{type(e).__name__}: {e}"}
        try:
            extras["motion_vs_ref_ext"] = motion_vs_reference_extension(dev)
        except Exception as e:
            extras["motion_vs_ref_ext"] = {"error": f"{type(e).__name__}: {e}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "steps_timed": steps_timed, "warmup": warmup,
        "ms_per_step": ms / steps_timed, "timed_region_s": ms / 1000, "host_ms_per_step": host_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1]: batch of {B} volumes/GPU/step, {shape[0]}^3 @0.5mm, label maps of the reference's bundled sub-sta21 / sub-sta30 / sub-sta38 (sample id k: subject k mod 3, own sub-class counts), deformation+GMM+gamma+bias+blur+resample+noise, all stage probs=1, Philox noise, ScaleIntensity fused",
                   "shape": list(shape), "batch_per_gpu": B, "coarse_fraction_f3": round(f3, 4),
                   "l2": f"every volume of a step has its own label map and intermediates ({B * nvox * 4 / 2**20:.0f} MiB per float buffer per step > 126 MB L2); the 3 subjects' packed inputs (48 MiB each) are shared across samples",
                   "host_affinity": numa},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_bytes[0], "d2h_bytes_per_step": e2e_bytes[1], "micro_batches_per_step": micro, "steps": e2e_steps},
        "e2e_packed_inputs": e2e_packed,
        "e2e_dataset_cache": e2e_cache,
        "host_dma": host_dma,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic_all.get(name), "peak_source": peak_src,
                     "per_kernel": per_kernel},
        "cpu_baseline": cpu,
    }
    line.update(extras)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--shape", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--depth", type=int, default=2, help="buffer slots of the host pipelines (e2e legs); 2 slots = 2.5 GiB of pinned host memory per rank at 256^3 / batch 8")
    ap.add_argument("--micro", type=int, default=4, help="micro-batches per step in the e2e legs (pipeline granularity)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs (profiling runs only)")
    ap.add_argument("--no-packed-e2e", action="store_true", help="skip the host-buffer leg with bit-packed seed inputs")
    ap.add_argument("--no-cache-e2e", action="store_true", help="skip the dataset-cache leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[2] artifact throughput and the motion-kernel comparison")
    args = ap.parse_args()
    shape = (args.shape,) * 3
    if args.impl == "reference":
        args.steps = min(args.steps, 3)
        run_reference(args, shape)
    else:
        run_ours(args, shape)


if __name__ == "__main__":
    main()
