#!/usr/bin/env python
"""Benchmark of the FetalSynthGen per-sample generation path (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels)
    python bench.py --impl reference --steps K --warmup W    # CPU port of the reference path

One *step* = one batch of `--batch` (default 8) synthetic 256^3 volumes per GPU through the whole
base pipeline (GMM -> warp+gamma+bias -> blur -> down-sample+noise -> up-sample /max), all
stage probabilities forced to 1 (fixed work), Philox noise.  `value` = volumes/s with inputs
resident in HBM; `e2e` = the same through the host-buffer API (`HostPipeline`: every step copies
its segmentation + 4 seed volumes from pinned host memory and its image + segmentation back, all
inside the timed region; copies of consecutive steps overlap with the kernels).  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
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


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU while the timed region runs: NVML polled every
    2 ms from a thread (a 25 ms timed region still yields ~10 samples); `nvidia-smi -lms` as a fallback."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.nvml, self._stop = index, [], None, None, False
        self.sm, self.mask, self.max_mhz = [], 0, None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES-relative index -> NVML handle through the PCI bus id
            import torch

            bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(torch.cuda.get_device_properties(self.index), "pci_bus_id") else None
            h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if int(pynvml.nvmlDeviceGetPciInfo(hi).bus) == int(bus):
                        h = hi
                        break
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml = (pynvml, h)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        pynvml, h = self.nvml
        while not self._stop:
            try:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.nvml is not None:
            self._stop = True
            time.sleep(0.005)
            reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.sm), "source": "nvml, 2 ms poll"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def bind_to_gpu_numa_node(local: int) -> str:
    """Pin this rank to the CPUs of its GPU's NUMA node before any pinned host memory is allocated:
    the pipeline's staging buffers then live next to the GPU's PCIe root (8 ranks move ~130 GB/s)."""
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


# ---------------------------------------------------------------------------------- synthetic inputs
def reference_motion_extension():
    """Baseline leg of `tools/bench_configs.py --config motion`: the reference's own slice-acquisition
    extension, built from the reference's sources into oracle/_ref by oracle/build_ref.py (None when it
    is absent).  It is only timed next to libfsg, never called by the product."""
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
    from fetalsyngen_b200.utils.phantom import label_phantom

    seg, seeds = label_phantom(shape)
    lab = sum(s.astype(np.int64) for s in seeds)
    rs = np.random.RandomState(seed)
    t = 0.0
    for _ in range(nvol):
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
def run_reference(args, shape):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
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
        "config": {"workload": f"base pipeline, all stage probs=1, {shape[0]}^3 @0.5mm phantom; one step = {workers} volumes (1 per host process)", "shape": list(shape)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": f"{args.steps} x {workers} volumes, numpy port of the reference path (oracle/np_oracle.py), one process per volume; host has {os.cpu_count()} logical CPUs (workers capped by min(cpus, 32, RAM / 6 GB))"},
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


ALGO_BYTES_PER_VOXEL = {  # SURVEY.md section 8(d); n = coarse-grid voxels / N
    "fsg_gmm": lambda r: 4 + 4,            # 4 seed bytes read + 4 B written (seed sum fused)
    "fsg_warp": lambda r: 10,              # img 4 + seg 1 read, img 4 + seg 1 written
    "fsg_blur3d": lambda r: 8,             # fused single pass (this round runs 3 passes = 24 B of real traffic)
    "fsg_resample": lambda r: 4 + 4 * r,
    "fsg_zoom": lambda r: 4 + 4 * r,
    "fsg_zoom_minmax": lambda r: 4 * r,
    "fsg_warp_shift": lambda r: 0,
}


def run_ours(args, shape):
    import torch
    import torch.distributed as dist

    from fetalsyngen_b200 import _lib
    from fetalsyngen_b200.utils.phantom import label_phantom

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

    seg_h, seeds_h = label_phantom(shape)
    gen = build_generator(shape, dev)
    eng = gen.engine(shape)
    seg_d = torch.from_numpy(seg_h).to(dev)
    seeds_d = [torch.from_numpy(s).to(dev) for s in seeds_h]
    out_img = torch.empty((B, *shape), dtype=torch.float32, device=dev)
    out_seg = torch.empty((B, *shape), dtype=torch.uint8, device=dev)

    from fetalsyngen_b200.sharding import step_ids

    counter = [0]

    def step():
        # rank r owns sample ids r, r+R, ...; every draw is a function of (1234, sample id), not of R
        ids = step_ids(counter[0], B, rank, world)
        counter[0] += 1
        gen.sample_batch([seg_d] * B, [seeds_d] * B, scale=True, out_img=out_img, out_seg=out_seg, sample_ids=ids, base_seed=1234)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.stats.reset()
    # CUDA events bracket the dominant kernel's launches inside the timed region (two event records
    # per step); bracketing every entry point costs ~0.4 ms of host time per step, so the full
    # per-kernel table comes from a few extra steps after the timed region
    _lib.stats.timing = {DOMINANT}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    _lib.stats.timing = False
    dominant = _lib.stats.elapsed_ms()
    launches = _lib.stats.total_kernels()  # kernels of libfsg launched inside the timed region
    clocks = sampler.stop() if rank == 0 else {}
    _lib.stats.reset()
    _lib.stats.timing = True
    for _ in range(3):
        step()
    _lib.stats.timing = False
    per_call = _lib.stats.elapsed_ms()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1000)

    # ---- e2e through the host-buffer API
    from fetalsyngen_b200.host_pipeline import HostPipeline

    e2e_steps = 0 if args.no_e2e else args.steps  # --no-e2e: profiling runs only
    # a step's B volumes travel as `micro` micro-batches through the pipeline (same bytes per step; the
    # pipeline fills / drains in units of B / micro volumes, so a short timed region is not dominated by the
    # first H2D and the last D2H)
    micro = args.micro if B % args.micro == 0 else 1
    mb = B // micro
    hp = HostPipeline(gen, mb, depth=(args.depth + 1) if e2e_steps else 1)
    hp.set_inputs([seg_h] * mb, [seeds_h] * mb)
    sink = [0.0]

    def consume(h_img, h_seg, params):
        sink[0] += float(h_img[0, 0, 0, 0]) + float(h_seg[-1, -1, -1, -1])  # the host reads the step's result

    if e2e_steps:
        hp.run((args.depth + 1) * micro, on_result=consume)
    barrier()
    t0 = time.perf_counter()
    hp.run(e2e_steps * micro, on_result=consume)  # returns when every step's image + segmentation is in host memory
    barrier()
    e2e_s = max(time.perf_counter() - t0, 1e-9)
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(t.item())

    # ---- same leg with the inputs in the bit-packed subject-cache format (FetalSynthDataset(packed_cache=...)):
    # uint8 segmentation + one uint16 word per voxel for all six sub-class counts instead of four int8 seed
    # volumes; the sub-class counts are drawn per sample and the label volume is unpacked on the device
    e2e_bytes = (hp.h2d_bytes * micro, hp.d2h_bytes * micro)
    e2e_packed = None
    if e2e_steps and not args.no_packed_e2e:
        from fetalsyngen_b200.data.packed import pack_seed_volumes

        del hp
        torch.cuda.empty_cache()
        per_count = {}
        for n in range(1, 7):
            _, sv = label_phantom(shape, n_sub=(n, n, n, n), seed=n)
            per_count[n] = {m + 1: sv[m] for m in range(4)}
        words, counts = pack_seed_volumes(per_count)
        del per_count
        hp = HostPipeline(gen, mb, depth=args.depth + 1, packed_counts=counts)
        hp.set_inputs_packed([seg_h] * mb, [words] * mb)
        hp.run((args.depth + 1) * micro, on_result=consume)
        barrier()
        t0 = time.perf_counter()
        hp.run(e2e_steps * micro, on_result=consume)
        barrier()
        t = torch.tensor([max(time.perf_counter() - t0, 1e-9)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_packed = {"value": world * B * e2e_steps / float(t.item()), "unit": UNIT, "h2d_bytes_per_step": hp.h2d_bytes * micro, "d2h_bytes_per_step": hp.d2h_bytes * micro,
                      "inputs": "uint8 segmentation + uint16 bit-packed seed words per voxel (subject-cache format), sub-class counts drawn per sample"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA-event time inside the timed region)
    peak, peak_src = peaks()
    name, (ncalls, tms) = DOMINANT, dominant[DOMINANT]
    if max(per_call.items(), key=lambda kv: kv[1][1] / kv[1][0])[0] != DOMINANT:
        print(f"bench.py: note: {DOMINANT} is no longer the slowest entry point: {per_call}", file=sys.stderr)
    r = 0.2  # mean coarse-grid fraction for spacing ~ U(0.5,1.5): E[(0.5/s)^3] ~ 0.2
    algo = ALGO_BYTES_PER_VOXEL.get(name, lambda r: 0)(r) * nvox * B
    achieved = algo / (tms / ncalls / 1000) / 1e9 if tms > 0 else 0.0
    traffic = None
    prof = ROOT / "profiles" / "dominant_kernel_traffic.json"
    if prof.exists():
        traffic = json.loads(prof.read_text()).get(name)

    # ---- CPU baseline: bounded sample of the same workload on the host cores
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, dt = cpu_port_throughput(shape, 1, 1)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"1 volume of the same workload, numpy port of the reference path (oracle/np_oracle.py), {dt:.1f} s, single process"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1]: batch of {B} volumes/GPU/step, {shape[0]}^3 @0.5mm phantom, deformation+GMM+gamma+bias+blur+resample+noise, all stage probs=1, Philox noise, ScaleIntensity fused", "shape": list(shape), "batch_per_gpu": B, "l2": f"inputs larger than L2 ({B * nvox * 4 / 2**20:.0f} MiB per buffer per step)", "host_affinity": numa},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_bytes[0], "d2h_bytes_per_step": e2e_bytes[1], "micro_batches_per_step": micro},
        "e2e_packed_inputs": e2e_packed,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "per_call_ms": {k: round(v[1] / v[0], 4) for k, v in per_call.items()}},
        "cpu_baseline": cpu,
    }
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
    ap.add_argument("--depth", type=int, default=2, help="buffer slots of the host pipeline (e2e leg); 2 slots = 2.5 GiB of pinned host memory per rank at 256^3 / batch 8")
    ap.add_argument("--micro", type=int, default=4, help="micro-batches per step in the e2e leg (pipeline granularity)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs only)")
    ap.add_argument("--no-packed-e2e", action="store_true", help="skip the additional host-buffer leg with bit-packed seed inputs")
    args = ap.parse_args()
    shape = (args.shape,) * 3
    if args.impl == "reference":
        args.steps = min(args.steps, 3)
        run_reference(args, shape)
    else:
        run_ours(args, shape)


if __name__ == "__main__":
    main()
