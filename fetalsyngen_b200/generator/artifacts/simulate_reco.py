"""Acquisition + reconstruction simulator behind ``SimulateMotion``.  Same classes, constructor
arguments, draw order and data-dictionary keys as the reference
(``fetalsyngen/generator/artifacts/simulate_reco.py``: ``Scanner`` ``:57-466``,
``PSFReconstructor`` ``:469-774``, ``PSFreconstruction`` ``:38-54``).

Host: all scalar draws (numpy global RNG, reference order) and the <= 250 slice transforms
(``svort.py``).  Device: libfsg K5-motion kernels (``csrc/motion.cu``) — slice acquisition with
a compact PSF tap list, per-slice sums, gamma / Rician noise / signal-void slice artifacts, the
PSF scatter reconstruction with equalisation, 3^3 smoothing and the merge with the clean volume.
Volumes are ``(D,H,W)`` = the tensor's three axes, W fastest, exactly as the reference passes
``output.view(1,1,*shape)`` to its extension.

``inject`` (not in the reference API) lets the parity tests pin the per-pixel random tensors:
keys ``noise1_<stack>`` / ``noise2_<stack>`` (full-size N(0,1) arrays of a stack), ``void_<stack>``
(dict idx / yc / xc / theta / a / A / sx), ``perm_misreg`` / ``perm_kept`` (the two randperm
results), plus the Perlin keys of ``StructNoise.perlin_weight``.
"""
from __future__ import annotations

import numpy as np
import torch

from ... import _lib
from ...artifact_ops import ArtifactOps, _stream, upload
from ...engine import engine_for
from . import svort
from .svort import RigidTransform
from .utils import ReconMergeParams

STAGE_RICIAN = 48
F32 = np.float32


def _dev_f32(arr, device):
    return upload(arr, F32, device)


def _dev_i32(arr, device):
    return upload(arr, np.int32, device)


_SIDE_STREAMS: dict = {}


def _side_stream(device) -> "torch.cuda.Stream":
    """Per-device stream for the cheap mask acquisitions whose read-back decides which slices survive:
    on the main stream that read-back would wait for the previous stack's PSF acquisition."""
    key = str(torch.device(device))
    s = _SIDE_STREAMS.get(key)
    if s is None:
        s = torch.cuda.Stream(device=device, priority=-1)  # its small kernels must not queue behind the main stream's grids
        _SIDE_STREAMS[key] = s
    return s


def volume_xpairs(vol):
    """(v[i], v[i+1]) per voxel of a (D,H,W) device volume: the gather format of the PSF acquisition
    (built once per volume; a Scanner acquires 2-6 stacks from it)."""
    v = vol.contiguous()
    pairs = torch.empty((v.numel(), 2), dtype=torch.float32, device=v.device)
    _lib.call("fsg_volume_xpairs", v.data_ptr(), pairs.data_ptr(), v.numel(), _stream())
    return pairs


def volume_xyquads(vol):
    """(v[i], v[i+1], v[i+W], v[i+W+1]) per voxel: two 16-byte loads per trilinear sample."""
    v = vol.contiguous()
    quads = torch.empty((v.numel(), 4), dtype=torch.float32, device=v.device)
    _lib.call("fsg_volume_xyquads", v.data_ptr(), quads.data_ptr(), v.numel(), int(v.shape[-1]), _stream())
    return quads


def slice_acquisition(mat, vol, psf, slice_shape, res_slice, out=None, pairs=None):
    """``svort.slice_acquisition(mat, vol, None, None, psf, slice_shape, res_slice, False, False)``
    (slice_acq.py:193-226): mat (n,3,4) host float32, vol (D,H,W) device float32, psf host array.
    ``pairs``: ``volume_xpairs(vol)`` — same result, half the gather instructions."""
    dev = vol.device
    taps, radius = svort.psf_taps(psf)
    n, (h, w) = mat.shape[0], slice_shape
    D, H, W = (int(s) for s in vol.shape[-3:])
    out = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev) if out is None else out
    t_d, taps_d = _dev_f32(mat, dev), _dev_f32(taps, dev)
    name, src = "fsg_slice_acq_forward", vol
    if pairs is not None:
        name, src = ("fsg_slice_acq_forward_xyquads" if pairs.shape[-1] == 4 else "fsg_slice_acq_forward_xpairs"), pairs
    _lib.call(name, t_d.data_ptr(), src.data_ptr(), taps_d.data_ptr(), int(taps.shape[0]), float(radius), out.data_ptr(), n, h, w, D, H, W, float(F32(res_slice)), _stream())
    return out


def slice_acquisition_adjoint(mat, psf, slices, vol_shape, res_slice, slice_idx=None, equalize=True):
    """``svort.slice_acquisition_adjoint(mat, psf, slices, None, None, vol_shape, res_slice, True, equalize)``
    (slice_acq.py:229-263).  ``slice_idx``: rows of ``slices`` that the n transforms refer to."""
    dev = slices.device
    taps, radius = svort.psf_taps(psf)
    radius += 1.0  # the scatter rounds to the nearest voxel
    n = mat.shape[0]
    h, w = (int(s) for s in slices.shape[-2:])
    D, H, W = (int(s) for s in vol_shape)
    vol = torch.empty((1, 1, D, H, W), dtype=torch.float32, device=dev)
    wgt = torch.empty((1, 1, D, H, W), dtype=torch.float32, device=dev)
    acc = torch.empty((D * H * W, 2), dtype=torch.float32, device=dev)  # interleaved (value, weight) accumulator
    t_d, taps_d, psf_d = _dev_f32(mat, dev), _dev_f32(taps, dev), _dev_f32(psf, dev)
    idx_d = None if slice_idx is None else _dev_i32(slice_idx, dev)
    dp, hp, wp = psf.shape
    _lib.call("fsg_slice_acq_adjoint", t_d.data_ptr(), psf_d.data_ptr(), dp, hp, wp, taps_d.data_ptr(), int(taps.shape[0]), float(radius), slices.data_ptr(),
              None if idx_d is None else idx_d.data_ptr(), vol.data_ptr(), wgt.data_ptr(), acc.data_ptr(), n, h, w, D, H, W, float(F32(res_slice)), int(bool(equalize)), _stream())
    return vol, wgt


class Scanner:
    """Simulates stacks of thick 2-D slices of a volume under inter-slice rigid motion
    (simulate_reco.py:57-466)."""

    def __init__(self, resolution_slice_fac_min, resolution_slice_fac_max, resolution_slice_max, slice_thickness_min, slice_thickness_max, gap_min, gap_max,
                 min_num_stack, max_num_stack, max_num_slices, noise_sigma_min, noise_sigma_max, TR_min, TR_max, prob_gamma, gamma_std, prob_void, slice_size,
                 restrict_transform: bool, txy: float, resolution_recon: float = None, slice_noise_threshold: float = 0.1):
        self.resolution_slice_fac_min = resolution_slice_fac_min
        self.resolution_slice_fac_max = resolution_slice_fac_max
        self.resolution_slice_max = resolution_slice_max
        self.slice_thickness_min = slice_thickness_min
        self.slice_thickness_max = slice_thickness_max
        self.gap_min = gap_min
        self.gap_max = gap_max
        self.min_num_stack = min_num_stack
        self.max_num_stack = max_num_stack
        self.max_num_slices = max_num_slices
        self.noise_sigma_min = noise_sigma_min
        self.noise_sigma_max = noise_sigma_max
        self.TR_min = TR_min
        self.TR_max = TR_max
        self.prob_gamma = prob_gamma
        self.gamma_std = gamma_std
        self.prob_void = prob_void
        self.slice_size = slice_size
        self.resolution_recon = resolution_recon
        self.restrict_transform = restrict_transform
        self.txy = txy
        self.slice_noise_threshold = slice_noise_threshold

    # ------------------------------------------------------------------ draws
    def get_resolution(self, data, genparams: dict = {}):
        resolution = data["resolution"]
        if "resolution_slice_fac" not in genparams:
            resolution_slice = np.random.uniform(self.resolution_slice_fac_min * resolution, min(self.resolution_slice_fac_max * resolution, self.resolution_slice_max))
        else:
            resolution_slice = genparams["resolution_slice_fac"]
        data["resolution_recon"] = self.resolution_recon if self.resolution_recon is not None else np.random.uniform(resolution, resolution_slice)
        data["resolution_slice"] = resolution_slice
        data["slice_thickness"] = np.random.uniform(self.slice_thickness_min, self.slice_thickness_max) if "slice_thickness" not in genparams else genparams["slice_thickness"]
        data["gap"] = np.random.uniform(self.gap_min, self.gap_max) if "gap" not in genparams else genparams["gap"]
        return data

    def sample_time(self, n_slice, genparams: dict = {}):
        TR = np.random.uniform(self.TR_min, self.TR_max) if "TR" not in genparams else genparams["TR"]
        return np.arange(n_slice) * TR

    # ------------------------------------------------------------------ slice artifacts (in place)
    def random_gamma(self, slices, genparams: dict = {}):
        if np.random.rand() < self.prob_gamma:
            gamma = np.exp(self.gamma_std * np.random.randn(1)[0]) if "gamma" not in genparams else genparams["gamma"]
            ws = torch.empty(1, dtype=torch.float32, device=slices.device)
            _lib.call("fsg_slice_gamma", slices.data_ptr(), slices.numel(), float(F32(gamma)), ws.data_ptr(), _stream())
        return slices

    def add_noise(self, slices, genparams: dict = {}, noise=None, rng=(0, 0)):
        sigma = np.random.uniform(self.noise_sigma_min, self.noise_sigma_max) if "noise_sigma" not in genparams else genparams["noise_sigma"]
        n1 = n2 = None
        if noise is not None:
            if any(int(np.asarray(v).size) != slices.numel() for v in noise):
                raise ValueError(f"injected noise must have one value per slice pixel ({slices.numel()})")
            n1, n2 = (_dev_f32(v, slices.device) for v in noise)
        _lib.call("fsg_slice_rician", slices.data_ptr(), slices.numel(), float(F32(self.slice_noise_threshold)), float(F32(sigma)),
                  None if n1 is None else n1.data_ptr(), None if n2 is None else n2.data_ptr(), _lib.Rng(rng[0] & (2**64 - 1), rng[1], STAGE_RICIAN, 0), _stream())
        return slices

    def signal_void(self, slices, void=None):
        """Draws of :273-288 from torch's CPU generator (the reference uses the device generator)."""
        n_all = slices.shape[0]
        if void is None:
            idx = torch.nonzero(torch.rand(n_all) < self.prob_void)[:, 0]
            n = int(idx.numel())
            if n == 0:
                return slices
            h, w = slices.shape[-2:]
            yc = (torch.rand(n) - 0.5) * (h - 1)
            xc = (torch.rand(n) - 0.5) * (w - 1)
            theta = 2 * np.pi * torch.rand(n)
            a = 30 + torch.rand(n) * 90
            A = torch.rand(n) * 0.5 + 0.5
            sx = torch.rand(n) * 30 + 39
            void = {"idx": idx.numpy(), "yc": yc.numpy(), "xc": xc.numpy(), "theta": theta.numpy(), "a": a.numpy(), "A": A.numpy(), "sx": sx.numpy()}
        n = len(void["idx"])
        if n == 0:
            return slices
        params = np.stack([np.asarray(void[k], dtype=F32).reshape(-1) for k in ("yc", "xc", "theta", "a", "A", "sx")], -1)
        p_d, i_d = _dev_f32(params, slices.device), _dev_i32(np.asarray(void["idx"]).reshape(-1), slices.device)
        _lib.call("fsg_slice_void", slices.data_ptr(), int(slices.shape[-2]), int(slices.shape[-1]), i_d.data_ptr(), p_d.data_ptr(), n, _stream())
        return slices

    # ------------------------------------------------------------------ acquisition
    def scan(self, data, genparams: dict = {}, inject: dict | None = None):
        inject = inject or {}
        data = self.get_resolution(data, genparams={})
        res, res_r, res_s = data["resolution"], data["resolution_recon"], data["resolution_slice"]
        s_thick, gap = data["slice_thickness"], data["gap"]
        vol = data["volume"]
        device = vol.device
        if res_r != res:
            # SimulateMotion pins resolution_recon to the volume resolution (artifacts.py:402), so the
            # grid_sample re-gridding of :320-331 is never reached on the generation path.
            raise NotImplementedError("Scanner.scan: resolution_recon != resolution is not on the generator path")
        data["volume_gt"], data["seg_gt"] = vol, data["seg"]
        psf_acq = svort.get_PSF(res_ratio=(res_s / res, res_s / res, s_thick / res))
        psf_rec = svort.get_PSF(res_ratio=(res_s / res_r, res_s / res_r, s_thick / res_r))
        psf_one = svort.get_PSF(0)
        data["psf_rec"], data["psf_acq"] = psf_rec, psf_acq
        vs = vol.shape
        if self.slice_size is None:
            ss = int(np.sqrt((vs[-1] ** 2 + vs[-2] ** 2 + vs[-3] ** 2) / 2.0) * res / res_s)
            ss = int(np.ceil(ss / 32.0) * 32)
        else:
            ss = self.slice_size
        ns = int(max(vs) * res / gap) + 2

        stacks, stacks_no_psf, transforms, transforms_gt, positions = [], [], [], [], []
        num_stacks = np.random.randint(self.min_num_stack, self.max_num_stack + 1)
        rng_seed = int(torch.randint(0, 2**62, (1,)).item())
        sums_d = torch.empty(ns, dtype=torch.float32, device=device)
        vol_pairs = volume_xyquads(vol)  # 2 x 16-byte gathers per trilinear sample instead of 8 x 4-byte: the acquisition is L1-wavefront bound
        main, side = torch.cuda.current_stream(device), _side_stream(device)
        side.wait_stream(main)  # the mask volume was produced on the main stream
        attempt = 0
        while True:
            transform_init = svort.random_init_stack_transforms(ns, gap, self.restrict_transform, self.txy)
            ts = self.sample_time(ns)
            transform_motion = svort.sample_motion(ts, True)
            interleave_idx = svort.interleave_index(ns, np.random.randint(2, int(np.sqrt(ns)) + 1))
            transform_motion = transform_motion[interleave_idx]
            transform_target = transform_motion.compose(transform_init)
            mat = svort.mat_update_resolution(transform_target.matrix(), res_r, res)
            # The reference acquires the image stack and the mask stack, then keeps the slices whose mask
            # content passes a random threshold (simulate_reco.py:355-371).  Slices are independent, so the
            # cheap one-tap mask stack goes first: the read-back that decides which slices survive waits for
            # it alone, and the PSF acquisition of the image runs for the surviving run of slices only.
            with torch.cuda.stream(side):
                slices_no_psf = slice_acquisition(mat, data["mask"], psf_one, (ss, ss), res_s / res)
                _lib.call("fsg_slice_sums", slices_no_psf.data_ptr(), ns, ss * ss, sums_d.data_ptr(), _stream())
                nnz = sums_d.cpu().numpy()  # synchronises the side stream only
            slices_no_psf.record_stream(main)  # consumed on the main stream later (PSFReconstructor)
            idx = nnz > (nnz.max() * np.random.uniform(0.1, 0.3))
            if idx.sum() == 0:
                continue
            nz = np.nonzero(idx)[0]
            idx[nz[0] : nz[-1]] = True
            lo, hi = int(nz[0]), int(nz[-1]) + 1  # the kept slices are one contiguous run
            slices = slice_acquisition(np.ascontiguousarray(mat[lo:hi]), vol, psf_acq, (ss, ss), res_s / res, pairs=vol_pairs)
            slices_no_psf = slices_no_psf[lo:hi]
            transform_init = svort.reset_transform(transform_init[idx])
            transform_target = transform_target[idx]
            k = attempt
            attempt += 1
            slices = self.random_gamma(slices)
            noise = (inject[f"noise1_{k}"], inject[f"noise2_{k}"]) if f"noise1_{k}" in inject else None
            slices = self.add_noise(slices, noise=noise, rng=(rng_seed, k))
            slices = self.signal_void(slices, inject.get(f"void_{k}"))
            if self.max_num_slices is not None and sum(st.shape[0] for st in stacks) + slices.shape[0] >= self.max_num_slices:
                break
            stacks.append(slices)
            stacks_no_psf.append(slices_no_psf)
            transforms.append(transform_init)
            transforms_gt.append(transform_target)
            positions.append(np.arange(slices.shape[0], dtype=F32) - slices.shape[0] // 2)
            if len(stacks) >= num_stacks:
                break
        main.wait_stream(side)  # the mask stacks are read on the main stream from here on
        stacks_ids = np.random.choice(20, len(stacks), replace=False)
        data["positions"] = np.concatenate([np.stack((positions[i], np.full_like(positions[i], s_i)), -1) for i, s_i in enumerate(stacks_ids)], 0)
        data["slice_shape"] = (ss, ss)
        data["volume_shape"] = tuple(int(s) for s in vs[-3:])
        data["stacks"] = torch.cat(stacks, 0)
        data["stacks_no_psf"] = torch.cat(stacks_no_psf, 0)
        transforms, transforms_gt = RigidTransform.cat(transforms), RigidTransform.cat(transforms_gt)
        data["transforms"], data["transforms_angle"] = transforms.matrix(), transforms
        data["transforms_gt"], data["transforms_gt_angle"] = transforms_gt.matrix(), transforms_gt
        data.pop("volume")
        return data


class PSFReconstructor:
    """PSF scatter reconstruction with randomised slice mis-registration, slice removal, smoothing
    and merge with the clean volume (simulate_reco.py:469-774)."""

    def __init__(self, prob_misreg_slice: float, slices_misreg_ratio: float, prob_misreg_stack: float, txy: float, prob_merge: float, merge_params: ReconMergeParams,
                 prob_smooth: float, prob_rm_slices: float, rm_slices_min: float, rm_slices_max: float):
        self.prob_misreg_slice = prob_misreg_slice
        self.slices_misreg_ratio = slices_misreg_ratio
        self.prob_misreg_stack = prob_misreg_stack
        self.txy_stack = txy
        self.prob_merge = prob_merge
        self.merge_params = merge_params
        assert merge_params.merge_type in ["gaussian", "perlin"], f"Merge type {merge_params.merge_type} not supported, only gaussian and perlin are supported."
        self.prob_smooth = prob_smooth
        self.prob_rm_slices = prob_rm_slices
        self.rm_slices_min = rm_slices_min
        self.rm_slices_max = rm_slices_max

    def sample_seeds(self, genparams: dict = {}):
        self._smooth_volume_on = np.random.rand() < self.prob_smooth
        self._rm_slices_on = np.random.rand() < self.prob_rm_slices
        self._misreg_slice_on = np.random.rand() < self.prob_misreg_slice
        if "rm_slices_ratio" in genparams:
            self._rm_slices_ratio = genparams["rm_slices_ratio"]
        else:
            self._rm_slices_ratio = np.random.uniform(self.rm_slices_min, self.rm_slices_max) if self._rm_slices_on else None
        self._misreg_stack_on = []
        self._merge_volume_on = np.random.rand() < self.prob_merge
        mp = self.merge_params
        if mp.merge_type == "gaussian":
            self._ngaussians_merge = genparams["ngaussians_merge"] if "ngaussians_merge" in genparams else np.random.randint(mp.gauss_ngaussians_min, mp.gauss_ngaussians_max)
        elif mp.merge_type == "perlin":
            self._res = genparams["res"] if "res" in genparams else np.random.choice(mp.perlin_res_list)
            self._octave = genparams["octave"] if "octave" in genparams else np.random.choice(mp.perlin_octaves_list)

    def get_seeds(self):
        seeds = {"smooth_volume_on": self._smooth_volume_on, "rm_slices_on": self._rm_slices_on, "rm_slices_ratio": self._rm_slices_ratio,
                 "misreg_stack_on": self._misreg_stack_on, "misreg_slice_on": self._misreg_slice_on, "merge_volume_on": self._merge_volume_on}
        if self.merge_params.merge_type == "gaussian":
            seeds["merge_type"] = "gaussian"
            seeds["ngaussians_merge"] = self._ngaussians_merge
        elif self.merge_params.merge_type == "perlin":
            seeds["merge_type"] = "perlin"
            seeds["res"] = self._res
            seeds["octave"] = self._octave
        return seeds

    # ------------------------------------------------------------------ transforms (host)
    def misregistration_trf(self, positions, base: RigidTransform) -> RigidTransform:
        nslices = len(positions)
        rand_angle = np.zeros((nslices, 6), dtype=F32)
        for pos in np.unique(positions[:, 1]):
            self._misreg_stack_on.append(np.random.rand() < self.prob_misreg_stack)
            if not self._misreg_stack_on[-1]:
                continue
            idx = np.where(positions[:, 1] == pos)[0]
            tx = np.ones(len(idx), dtype=F32) * np.random.uniform(-self.txy_stack, self.txy_stack)
            ty = np.ones(len(idx), dtype=F32) * np.random.uniform(-self.txy_stack, self.txy_stack)
            rand_angle[idx, 3:] = svort.random_angle(len(idx), restricted=True)
            rand_angle[idx, :3] = np.stack((tx, ty, np.zeros_like(tx)), -1)
        return RigidTransform(rand_angle, trans_first=True).compose(base)

    def misregister_slices(self, trf: RigidTransform, trf_gt: RigidTransform, perm=None) -> RigidTransform:
        trf1, trf2 = trf.axisangle(), trf_gt.axisangle()
        if self._misreg_slice_on:
            perm = torch.randperm(trf2.shape[0]).numpy() if perm is None else np.asarray(perm)
            idx_misreg = perm[: int(self.slices_misreg_ratio * trf2.shape[0])][:1]
            trf2[idx_misreg] = trf1[idx_misreg]
        return RigidTransform(trf2, trans_first=True)

    def kept_slices_idx(self, nslices, perm=None):
        if self._rm_slices_on:
            n = int(nslices * self._rm_slices_ratio)
            perm = torch.randperm(nslices).numpy() if perm is None else np.asarray(perm)
            return perm[n:]
        return np.arange(nslices)

    # ------------------------------------------------------------------ reconstruction (device)
    def merge_weight(self, eng, ops, mask_u8, out, minmax, inject=None):
        """Raw merge weight into ``out``; returns (increase, use_minmax, keep-alive)."""
        mp = self.merge_params
        if mp.merge_type == "perlin":
            from ..augmentation.artifacts import StructNoise

            keep = StructNoise.perlin_weight(eng, ops, int(self._res), int(self._octave), mp.perlin_persistence, mp.perlin_lacunarity, out, minmax, inject)
            return float(mp.perlin_increase_size), True, keep
        from ..augmentation.artifacts import _rng_pair

        centers, count = ops.sample_voxels(mask_u8, int(self._ngaussians_merge), match=1, transpose_out=True, rng=_rng_pair())
        sig = torch.cat([torch.clamp(20 + 10 * torch.randn(1), 5, 40) for _ in range(count)]).float() if count else torch.zeros(0)
        ops.mog(centers[:count].contiguous(), sig[:, None].repeat(1, 3).contiguous().to(eng.device), out=out)
        return 0.0, False, (centers, sig)

    def recon_psf(self, data, inject: dict | None = None):
        inject = inject or {}
        self.sample_seeds()
        stacks = data["stacks"]
        device = stacks.device
        trf = self.misregister_slices(data["transforms_angle"], data["transforms_gt_angle"], inject.get("perm_misreg"))
        trf = self.misregistration_trf(data["positions"], trf)
        kept_idx = self.kept_slices_idx(stacks.shape[0], inject.get("perm_kept"))
        res_slice = data["resolution_slice"] / data["resolution_recon"]
        D, H, W = data["volume_shape"]
        volume, _ = slice_acquisition_adjoint(trf.matrix()[kept_idx], data["psf_rec"], stacks, (D, H, W), res_slice, slice_idx=kept_idx, equalize=True)
        eng = engine_for(device, (D, H, W), (1.0, 1.0, 1.0))
        ops = ArtifactOps(eng)
        out = torch.empty((D, H, W), dtype=torch.float32, device=device)
        gt = data["volume_gt"].contiguous()
        weight = None
        if self._merge_volume_on:
            weight = ops.f32("b")
            minmax = torch.zeros(2, dtype=torch.float32, device=device)
            mask_u8 = eng.to_u8((data["seg_gt"] > 0).float().view(-1))
            increase, use_mm, keep = self.merge_weight(eng, ops, mask_u8, weight, minmax, inject)
            _lib.call("fsg_recon_merge", volume.data_ptr(), gt.data_ptr(), weight.data_ptr(), minmax.data_ptr() if use_mm else None, increase,
                      int(bool(self._smooth_volume_on)), D, H, W, out.data_ptr(), _stream())
        elif self._smooth_volume_on:
            _lib.call("fsg_recon_merge", volume.data_ptr(), None, None, None, 0.0, 1, D, H, W, out.data_ptr(), _stream())
        else:
            out = volume.view(D, H, W)
        return out.view(1, 1, D, H, W), weight


def simulate_motion(art, output, seg, resolution, inject: dict | None = None):
    """Body of ``SimulateMotion.__call__`` after its gate (artifacts.py:389-421)."""
    from dataclasses import asdict, fields

    device = output.device
    if device.type != "cuda":
        raise _lib.FsgError("SimulateMotion runs on CUDA tensors only; there is no CPU fallback")
    shape = tuple(int(s) for s in output.shape[-3:])
    res_ = np.float64(resolution[0])
    eng = engine_for(device, shape, (1.0, 1.0, 1.0))
    vol = output.to(torch.float32).contiguous().view(shape)
    segf = seg.to(device)
    d = {
        "resolution": res_,
        "volume": vol,
        "mask": (segf > 0).to(torch.float32).contiguous().view(shape),
        "seg": segf.to(torch.float32).contiguous().view(shape),
        "affine": torch.diag(torch.tensor(list(resolution) + [1])).to(device),
        "threshold": 0.1,
    }
    art.scanner_args.resolution_recon = res_
    scanner = Scanner(**asdict(art.scanner_args))
    d_scan = scanner.scan(d, inject=inject)
    recon = PSFReconstructor(**{f.name: getattr(art.recon_args, f.name) for f in fields(art.recon_args)})
    out, _ = recon.recon_psf(d_scan, inject=inject)
    metadata = {
        "resolution_recon": d_scan["resolution_recon"],
        "resolution_slice": d_scan["resolution_slice"],
        "slice_thickness": d_scan["slice_thickness"],
        "gap": d_scan["gap"],
        "nstacks": len(np.unique(d_scan["positions"][:, 1])),
    }
    metadata.update(recon.get_seeds())
    return out.squeeze(), metadata
