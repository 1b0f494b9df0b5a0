"""SR-artifact augmentations.  Same class names, constructors, call signatures and metadata as
the reference (``fetalsyngen/generator/augmentation/artifacts.py``): every ``__call__`` is
``(output, seg, device, genparams={}, **kwargs) -> (output, metadata)``.

Host side: the scalar draws, in the reference's numpy order.  Device side: libfsg K5 kernels
(``csrc/artifacts.cu``) — MoG with per-row blob culling, exponential-race voxel sampling,
fractal Perlin, separable blur, exact binary morphology as distance transforms.  ``inject``
(not in the reference API) feeds the reference's captured random tensors for the parity tests.
"""
from __future__ import annotations

import numpy as np
import torch

from ... import _lib
from ...artifact_ops import STAGE_PYRAMID, STAGE_RING, ArtifactOps, _stream, upload
from ...engine import SamplePlan, engine_for
from ..artifacts.utils import ReconParams, ScannerParams, StructNoiseMergeParams
from .synthseg import RandTransform


def _ops(output, device, resolution=(1.0, 1.0, 1.0)):
    eng = engine_for(device if str(device).startswith("cuda") else output.device, tuple(output.shape), (1.0, 1.0, 1.0))
    return eng, ArtifactOps(eng)


def _seg_u8(eng, seg):
    return eng.to_u8(seg.to(eng.device)).contiguous().view(-1)


def _rng_pair():
    return int(torch.randint(0, 2**62, (1,)).item()), int(torch.randint(0, 2**31, (1,)).item())


class BlurCortex(RandTransform):
    """Blurs the cortex at random Gaussian blobs (artifacts.py:24-133)."""

    def __init__(self, prob: float, cortex_label: int, nblur_min: int, nblur_max: int, sigma_gamma_loc: int = 3, sigma_gamma_scale: int = 1,
                 std_blur_shape: int = 2, std_blur_scale: int = 1):
        self.prob = prob
        self.cortex_label = cortex_label
        self.nblur_min = nblur_min
        self.nblur_max = nblur_max
        self.sigma_gamma_loc = sigma_gamma_loc
        self.sigma_gamma_scale = sigma_gamma_scale
        self.std_blur_shape = std_blur_shape
        self.std_blur_scale = std_blur_scale

    def __call__(self, output, seg, device, genparams: dict = {}, inject: dict | None = None, **kwargs):
        if not (np.random.rand() < self.prob or len(genparams.keys()) > 0):
            return output, {"nblur": None}
        inject = inject or {}
        nblur = np.random.randint(self.nblur_min, self.nblur_max) if "nblur" not in genparams.keys() else genparams["nblur"]
        std_blurs = np.random.gamma(self.std_blur_shape, self.std_blur_scale, 3)
        if "std_blurs" in inject:
            nblur, std_blurs = int(inject["nblur"]), np.asarray(inject["std_blurs"], dtype=np.float64)
        eng, ops = _ops(output, device)
        x, y, z = eng.shape
        seg8 = _seg_u8(eng, seg)
        src = output.to(eng.device, torch.float32).contiguous().view(-1)
        if "centers" in inject:  # reference voxel indices (i0, i1, i2): blob at axis position (i2, i1, i0)
            c = np.asarray(inject["centers"], dtype=np.float32)[:, ::-1].copy()
            centers, count = upload(c, np.float32, eng.device), c.shape[0]
        else:
            # frontal-lobe prior (blur_proba, :63-81): centres (0,y,z//2),(x,y,z//2) unpacked x<->last axis
            prior = ([(z // 2, y, 0), (z // 2, y, x)], [[x // 5] * 3, [y // 5] * 3])
            centers, count = ops.sample_voxels(seg8, int(nblur), match=self.cortex_label, prior=prior, transpose_out=True, rng=_rng_pair())
        sig = np.asarray(inject["sigmas"], dtype=np.float64) if "sigmas" in inject else np.random.gamma(self.sigma_gamma_loc, self.sigma_gamma_scale, (int(nblur), 3))
        sig_axis = upload(sig[:count, ::-1], np.float32, eng.device)
        blurred, t1, t2 = ops.f32("a"), ops.f32("b"), ops.f32("c")
        eng.sepconv([SamplePlan(stds=std_blurs)], [src], [blurred], [t1], [t2], positions=False)
        out = torch.empty_like(src)
        ops.mog(centers[:count].contiguous(), sig_axis, blend=(src, blurred, out))
        return out.view(eng.shape), {"nblur": nblur}


class StructNoise(RandTransform):
    """Multi-scale structured noise merged by a Perlin (or MoG) weight inside the brain
    (artifacts.py:136-342)."""

    def __init__(self, prob: float, wm_label: int, std_min: float, std_max: float, merge_params: StructNoiseMergeParams, nstages_min: int = 1, nstages_max: int = 5):
        self.prob = prob
        self.wm_label = wm_label
        self.nstages_min = nstages_min
        self.nstages_max = nstages_max
        self.std_min = std_min
        self.std_max = std_max
        self.merge_params = merge_params

    def sample_seeds(self, genparams: dict = {}):
        self.nstages = np.random.randint(self.nstages_min, self.nstages_max) if "nstages" not in genparams else genparams["nstages"]
        self.noise_std = self.std_min + (self.std_max - self.std_min) * np.random.rand()
        mp = self.merge_params
        if mp.merge_type == "gaussian":
            self.gauss_nloc = np.random.randint(mp.gauss_nloc_min, mp.gauss_nloc_max) if "nloc" not in genparams else genparams["nloc"]
        elif mp.merge_type == "perlin":
            self._res = genparams["res"] if "res" in genparams else np.random.choice(mp.perlin_res_list)
            self._octave = genparams["octave"] if "octave" in genparams else np.random.choice(mp.perlin_octaves_list)

    def get_seeds(self):
        seeds = {"nstages": self.nstages, "noise_std": self.noise_std}
        if self.merge_params.merge_type == "gaussian":
            seeds["nloc"] = self.gauss_nloc
        elif self.merge_params.merge_type == "perlin":
            seeds["res"] = self._res
            seeds["octave"] = self._octave
        return seeds

    # ------------------------------------------------------------------ device pieces
    @staticmethod
    def perlin_weight(eng, ops, res, octaves, persistence, lacunarity, out, minmax, inject=None):
        """Raw fractal noise + min/max (generate_fractal_noise_3d, artifacts/utils.py:330-388)."""
        inject = inject or {}
        octs = (_lib.PerlinOctave * int(octaves))()
        keep = []
        freq, amp = 1, 1.0
        for o in range(int(octaves)):
            r = [int(freq * res)] * 3
            if f"theta_{o}" in inject:
                u1, u2 = torch.from_numpy(inject[f"theta_{o}"]), torch.from_numpy(inject[f"phi_{o}"])
            else:
                u1, u2 = torch.rand(r[0] + 1, r[1] + 1, r[2] + 1), torch.rand(r[0] + 1, r[1] + 1, r[2] + 1)
            theta, phi = 2 * torch.pi * u1, 2 * torch.pi * u2
            g = torch.stack((torch.sin(phi) * torch.cos(theta), torch.sin(phi) * torch.sin(theta), torch.cos(phi)), dim=-1)
            g[-1, :, :] = g[0, :, :]
            g[:, -1, :] = g[:, 0, :]
            g[:, :, -1] = g[:, :, 0]
            gd = upload(g.float().numpy(), np.float32, eng.device)  # non-blocking: a plain .to() here waits for the whole reconstruction
            keep.append(gd)
            octs[o].grad = gd.data_ptr()
            for a in range(3):
                lin = eng.tables._put(("linspace", r[a], eng.shape[a]), lambda a=a: torch.linspace(0, r[a], eng.shape[a]).numpy())
                octs[o].lin[a] = lin.data_ptr()
                octs[o].res[a] = r[a]
            octs[o].amp = float(amp)
            freq *= lacunarity
            amp *= persistence
        _lib.call("fsg_perlin", octs, int(octaves), *eng.shape, out.data_ptr(), minmax.data_ptr(), _stream())
        return keep

    def multiscale_noise(self, eng, ops, nstages, out, minmax, inject=None):
        """Noise pyramid of artifacts.py:308-320 into ``out`` (full resolution) + its min/max."""
        inject = inject or {}
        shape = eng.shape
        cur_shape = [s // 2**nstages for s in shape]
        a, b = ops.f32("pyr0"), ops.f32("pyr1")
        a[: int(np.prod(cur_shape))].zero_()
        seed, sid = _rng_pair()
        for k in range(nstages):
            nxt = [s // 2 ** (nstages - 1 - k) for s in shape]
            nz = None
            if f"randn_{k}" in inject:
                nz = torch.from_numpy(inject[f"randn_{k}"]).to(eng.device).contiguous().view(-1)
            ops.add_noise_noclamp(a, int(np.prod(cur_shape)), _lib.Rng(seed, sid, STAGE_PYRAMID + k, 0), nz)
            tabs = ops.upsample_tabs(cur_shape, nxt)
            last = k == nstages - 1
            if last:
                ops.zoom_tabs(a, cur_shape, tabs, nxt, minmax=minmax, reduce_only=True, post=1)
            ops.zoom_tabs(a, cur_shape, tabs, nxt, dst=out if last else b)
            if not last:
                a, b = b, a
            cur_shape = nxt

    def __call__(self, output, seg, device, genparams: dict = {}, inject: dict | None = None, **kwargs):
        if not (np.random.rand() < self.prob or "nloc" in genparams.keys()):
            return output, {}
        inject = inject or {}
        self.sample_seeds()
        if "nstages" in inject:
            self.nstages, self.noise_std = int(inject["nstages"]), float(inject["noise_std"])
            self._res, self._octave = int(inject["res"]), int(inject["octave"])
        eng, ops = _ops(output, device)
        seg8 = _seg_u8(eng, seg)
        src = output.to(eng.device, torch.float32).contiguous().view(-1)
        scal = torch.zeros(8, dtype=torch.float32, device=eng.device)
        lr, weight = ops.f32("a"), ops.f32("b")
        self.multiscale_noise(eng, ops, int(self.nstages), lr, scal[0:2], inject)
        _lib.call("fsg_minmax", src.data_ptr(), src.numel(), scal[2:4].data_ptr(), _stream())
        mp = self.merge_params
        increase = 0.0
        if mp.merge_type == "perlin":
            keep = self.perlin_weight(eng, ops, int(self._res), int(self._octave), mp.perlin_persistence, mp.perlin_lacunarity, weight, scal[4:6], inject)
            increase = float(mp.perlin_increase_size)
        elif mp.merge_type == "gaussian":
            centers, count = ops.sample_voxels(seg8, int(self.gauss_nloc), match=self.wm_label, transpose_out=True, rng=_rng_pair())
            sig = torch.clamp(mp.gauss_sigma_mu + mp.gauss_sigma_std * torch.randn(count), 1, 40).float()
            ops.mog(centers[:count].contiguous(), sig[:, None].repeat(1, 3).contiguous().to(eng.device), out=weight)
            scal[4:6] = torch.tensor([0.0, 1.0], device=eng.device)
        else:
            raise RuntimeError
        out = torch.empty_like(src)
        _lib.call("fsg_struct_blend", src.data_ptr(), seg8.data_ptr(), lr.data_ptr(), weight.data_ptr(), scal.data_ptr(), float(self.noise_std), increase,
                  out.data_ptr(), src.numel(), _stream())
        return out.view(eng.shape), self.get_seeds()


class SimulatedBoundaries(RandTransform):
    """No mask / halo / fuzzy brain-mask boundaries (artifacts.py:428-604)."""

    def __init__(self, prob_no_mask: float, prob_if_mask_halo: float, prob_if_mask_fuzzy: float):
        self.prob_no_mask = prob_no_mask
        self.prob_halo = prob_if_mask_halo
        self.prob_fuzzy = prob_if_mask_fuzzy
        self.reset_seeds()

    def reset_seeds(self):
        self.no_mask_on = None
        self.halo_on = None
        self.halo_radius = None
        self.fuzzy_on = None
        self.n_generate_fuzzy = None
        self.n_centers = None
        self.base_sigma = None

    def sample_seeds(self):
        self.reset_seeds()
        self.no_mask_on = np.random.rand() < self.prob_no_mask
        if not self.no_mask_on:
            self.halo_on = np.random.rand() < self.prob_halo
            if self.halo_on:
                self.halo_radius = np.random.randint(5, 15)
            self.fuzzy_on = np.random.rand() < self.prob_fuzzy
            if self.fuzzy_on:
                self.n_generate_fuzzy = np.random.randint(2, 5)
                self.n_centers = np.random.poisson(100)
                self.base_sigma = np.random.poisson(8)

    # ------------------------------------------------------------------ device pieces
    def build_halo(self, ops, mask, radius, out):
        """Dilation by ball(radius) == squared Euclidean distance <= radius^2 (:484-499)."""
        d, t = ops.u16("d0"), ops.u16("d1")
        ops.dist(mask, d, t, int(radius), 0)
        return ops.thresh(d, out, int(radius) ** 2)

    def generate_fuzzy_boundaries(self, ops, mask, out, keep=None, rng=(0, 0), it=0):
        """One fuzzy round (:501-522): 7^3 ring -> keep 10 % -> 3^3 count > 3 -> 5^3 closing."""
        n = ops.n
        t0, t1, t2 = ops.u8("m0"), ops.u8("m1"), ops.u8("m2")
        ops.box(mask, t0, t1, 7, 0)                                   # t0 = dilate(mask, 7)
        _lib.call("fsg_morph_ring", t0.data_ptr(), mask.data_ptr(), None if keep is None else keep.data_ptr(),
                  _lib.Rng(rng[0] & (2**64 - 1), rng[1], STAGE_RING + it, 0), 0.1, t2.data_ptr(), n, _stream())
        ops.box(t2, t0, t1, 3, 2)                                     # t0 = 3^3 count of the kept ring
        _lib.call("fsg_morph_count_merge", t0.data_ptr(), mask.data_ptr(), 3, t2.data_ptr(), n, _stream())
        ops.box(t2, t0, t1, 5, 0)                                     # closing: dilate 5 ...
        ops.box(t0, out, t1, 5, 1)                                    # ... erode 5
        return out

    def __call__(self, output, seg, device, genparams: dict = {}, inject: dict | None = None, **kwargs):
        inject = inject or {}
        self.sample_seeds()
        if "halo_radius" in inject:
            self.no_mask_on, self.halo_on, self.fuzzy_on = False, True, True
            self.halo_radius, self.n_generate_fuzzy = int(inject["halo_radius"]), int(inject["n_generate_fuzzy"])
            self.n_centers, self.base_sigma = int(inject["n_centers"]), int(inject["base_sigma"])
        metadata = {"no_mask_on": self.no_mask_on, "halo_on": self.halo_on, "fuzzy_on": self.fuzzy_on}
        if self.no_mask_on:
            return output, metadata
        eng, ops = _ops(output, seg.device if str(seg.device).startswith("cuda") else device)
        n = ops.n
        seg8 = _seg_u8(eng, seg)
        src = output.to(eng.device, torch.float32).contiguous().view(-1)
        mask = ops.u8("mask")
        _lib.call("fsg_label_mask", seg8.data_ptr(), -1, mask.data_ptr(), n, _stream())
        if self.halo_on:
            halo = ops.u8("halo")
            self.build_halo(ops, mask, self.halo_radius, halo)
            mask = halo
        out = torch.empty_like(src)
        if not self.fuzzy_on:
            _lib.call("fsg_mask_mul", src.data_ptr(), mask.data_ptr(), out.data_ptr(), n, _stream())
            return out.view(eng.shape), metadata
        rng = _rng_pair()
        cur, nxt = mask, ops.u8("fz0")
        spare = ops.u8("fz1")
        for it in range(self.n_generate_fuzzy):
            keep = None
            if f"keep_{it}" in inject:
                keep = torch.from_numpy(np.ascontiguousarray(inject[f"keep_{it}"], dtype=np.uint8)).to(eng.device).view(-1)
            self.generate_fuzzy_boundaries(ops, cur, nxt, keep, rng, it)
            cur, nxt = nxt, (spare if nxt is not spare else ops.u8("fz0"))
        mask_modif = cur
        if "centers" in inject:
            c = np.asarray(inject["centers"], dtype=np.float32)[:, ::-1].copy()
            centers, count = upload(c, np.float32, eng.device), c.shape[0]
        else:
            centers, count = ops.sample_voxels(mask, int(self.n_centers), labels2=mask_modif, transpose_out=True, rng=_rng_pair())
        sigmas = np.asarray(inject["sigmas"], dtype=np.float64) if "sigmas" in inject else np.array([self.base_sigma + 10 * np.random.beta(2, 5) for _ in range(count)])
        sig_axis = upload(np.repeat(sigmas[:count, None], 3, 1), np.float32, eng.device)
        mog = ops.f32("a")
        ops.mog(centers[:count].contiguous(), sig_axis, out=mog)
        n_dilate = 6 * (self.n_generate_fuzzy - 1)
        l1, t16 = ops.u16("d0"), ops.u16("d1")
        ops.dist(mask, l1, t16, max(n_dilate - 2, 1), 1)
        _lib.call("fsg_boundary_select", src.data_ptr(), mask.data_ptr(), mask_modif.data_ptr(), l1.data_ptr(), mog.data_ptr(), int(n_dilate), out.data_ptr(), None, n, _stream())
        return out.view(eng.shape), metadata


class SimulateMotion(RandTransform):
    """Slice acquisition with inter-slice motion + PSF reconstruction (artifacts.py:345-425)."""

    def __init__(self, prob: float, scanner_params: ScannerParams, recon_params: ReconParams):
        self.scanner_args = scanner_params
        self.recon_args = recon_params
        self.prob = prob

    def __call__(self, output, seg, device, genparams: dict = {}, **kwargs):
        if not (np.random.rand() < self.prob):
            return output, {}
        from ..artifacts.simulate_reco import simulate_motion

        return simulate_motion(self, output, seg, kwargs["resolution"])
