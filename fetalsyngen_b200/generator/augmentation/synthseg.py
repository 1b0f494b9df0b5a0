"""SynthSeg-style intensity augmentations.  Same class names, constructors, call signatures and
parameter dictionaries as the reference (``fetalsyngen/generator/augmentation/synthseg.py``).

Every class has two faces:
  * ``draw(...)``  — host-side parameter draws in the reference's numpy/torch RNG order; used by
    ``FetalSynthGen`` to build a fused launch plan (gamma and bias ride in the warp kernel's
    epilogue, noise in the down-sampling kernel, /max in the up-sampling kernel);
  * ``__call__``   — the reference's standalone stage contract, executed by the same kernels.
"""
from __future__ import annotations

import numpy as np
import torch

from ...tables import resample_stds


class RandTransform:
    def __call__(self, *args, **kwargs):
        raise NotImplementedError

    def random_uniform(self, min_val, max_val):
        return np.random.uniform(min_val, max_val)


def _engine(device, shape, resolution=(1.0, 1.0, 1.0)):
    from ...engine import engine_for

    return engine_for(device, tuple(shape), resolution)


class RandResample(RandTransform):
    """Resolution simulation: Gaussian blur + trilinear down-sampling to a random isotropic
    spacing, and the matching ``resize_back`` (synthseg.py:26-114)."""

    def __init__(self, prob: float, min_resolution: float, max_resolution: float):
        self.prob = prob
        self.min_resolution = min_resolution
        self.max_resolution = max_resolution

    def draw(self, input_resolution, genparams: dict = {}, inject: dict | None = None):
        """Returns (spacing float64[3] | None, stds float64[3] | None)   (synthseg.py:63-80)."""
        if not (np.random.rand() < self.prob or "spacing" in genparams.keys()):
            return None, None
        spacing = np.array([1.0, 1.0, 1.0]) * self.random_uniform(self.min_resolution, self.max_resolution) if "spacing" not in genparams.keys() else genparams["spacing"]
        spacing = np.array(spacing, dtype=np.float64)
        blur_u = np.random.rand()
        if inject and "stds" in inject:
            stds = np.asarray(inject["stds"], dtype=np.float64)
        else:
            stds = resample_stds(spacing, input_resolution, blur_u)
        return spacing, stds

    def __call__(self, output, input_resolution, device, genparams: dict = {}, inject: dict | None = None):
        from ...engine import SamplePlan

        spacing, stds = self.draw(input_resolution, genparams, inject)
        if spacing is None:
            return output, None, {"spacing": None}
        eng = _engine(device, output.shape, input_resolution)
        src = output.to(eng.device, torch.float32).contiguous().view(1, -1)
        blurred = torch.empty_like(src)
        tmp = torch.empty_like(src)
        eng.blur([stds], src, blurred, tmp)
        plan = SamplePlan(spacing=spacing, stds=stds)
        n = eng.lowres_shape(spacing)
        low = torch.empty((1, int(np.prod(n))), dtype=torch.float32, device=eng.device)
        info = eng.resample([plan], blurred, low)
        self._last_shape = tuple(output.shape)
        return low.view(n), info[0][1], {"spacing": spacing.tolist()}

    def resize_back(self, output_resized, factors):
        if factors is None:
            return output_resized
        factors = np.asarray(factors, dtype=np.float64)
        n = tuple(output_resized.shape)
        shape = tuple(int(np.round(n[a] * (1 / factors[a]))) for a in range(3))
        eng = _engine(output_resized.device, shape)
        src = output_resized.to(torch.float32).contiguous()
        dst = torch.empty((1, int(np.prod(shape))), dtype=torch.float32, device=eng.device)
        eng.zoom([src.view(-1)], [n], [1 / factors], dst, post=1)
        return dst.view(shape)


class RandBiasField(RandTransform):
    """Multiplicative smooth bias field exp(zoom(N(0, std) control grid)) (synthseg.py:117-188)."""

    def __init__(self, prob: float, scale_min: float, scale_max: float, std_min: float, std_max: float):
        self.prob = prob
        self.scale_min = scale_min
        self.scale_max = scale_max
        self.std_min = std_min
        self.std_max = std_max

    def draw(self, image_size, genparams: dict = {}, inject: dict | None = None, device_grids: bool = False):
        """Returns (bf_low float32 grid | None, params dict)   (synthseg.py:157-176).  device_grids: return
        ((size, std), params) instead; the grid is then drawn on the device (fsg_draw_grids)."""
        if not (np.random.rand() < self.prob or len(genparams.keys()) > 0):
            return None, {"bf_scale": None, "bf_std": None, "bf_size": None}
        bf_scale = self.scale_min + np.random.rand(1) * (self.scale_max - self.scale_min) if "bf_scale" not in genparams.keys() else genparams["bf_scale"]
        bf_size = np.round(bf_scale * np.array(image_size)).astype(int)
        bf_size = np.maximum(bf_size, 1).tolist()
        bf_std = self.std_min + (self.std_max - self.std_min) * np.random.rand(1) if "bf_std" not in genparams.keys() else genparams["bf_std"]
        if device_grids and not (inject and "bf_n" in inject):
            return (tuple(int(v) for v in bf_size), float(np.asarray(bf_std, dtype=np.float32).reshape(-1)[0])), {"bf_scale": bf_scale, "bf_std": bf_std, "bf_size": bf_size}
        if inject and "bf_n" in inject:
            n = np.asarray(inject["bf_n"], dtype=np.float32)
        else:
            n = torch.randn(bf_size, dtype=torch.float).numpy()
        bf_low = (np.asarray(bf_std, dtype=np.float32) * n).astype(np.float32)
        return bf_low, {"bf_scale": bf_scale, "bf_std": bf_std, "bf_size": bf_size}

    def __call__(self, output, device, genparams: dict = {}, inject: dict | None = None):
        from ...engine import SamplePlan

        bf_low, params = self.draw(output.shape, genparams, inject)
        if bf_low is None:
            return output, params
        eng = _engine(device, output.shape)
        src = output.to(eng.device, torch.float32).contiguous()
        dst = torch.empty_like(src)
        eng.warp([SamplePlan(bf_low=bf_low)], [src.view(-1)], None, [dst.view(-1)], None)
        return dst, params


class RandNoise(RandTransform):
    """Additive Gaussian noise, clamped at 0 (synthseg.py:191-235)."""

    def __init__(self, prob: float, std_min: float, std_max: float):
        self.prob = prob
        self.std_min = std_min
        self.std_max = std_max

    def draw(self, genparams: dict = {}):
        if not (np.random.rand() < self.prob or "noise_std" in genparams.keys()):
            return None
        noise_std = self.std_min + (self.std_max - self.std_min) * np.random.rand(1) if "noise_std" not in genparams.keys() else genparams["noise_std"]
        return float(np.asarray(noise_std, dtype=np.float32).reshape(-1)[0])

    def __call__(self, output, device, genparams: dict = {}, inject: dict | None = None):
        from ...engine import SamplePlan

        noise_std = self.draw(genparams)
        if noise_std is None:
            return output, {"noise_std": None}
        eng = _engine(device, output.shape)
        src = output.to(eng.device, torch.float32).contiguous()
        dst = torch.empty_like(src)
        plan = SamplePlan(noise_std=noise_std, rng_seed=int(torch.randint(0, 2**62, (1,)).item()))
        if inject and "noise" in inject:
            plan.noise = torch.as_tensor(inject["noise"], dtype=torch.float32).to(eng.device).contiguous().view(-1)
        eng.add_noise([plan], [src.view(-1)], [dst.view(-1)], numel=src.numel())
        return dst, {"noise_std": noise_std}


class RandGamma(RandTransform):
    """Gamma contrast 300*(x/300)^gamma, gamma = exp(std*N) (synthseg.py:238-275)."""

    def __init__(self, prob: float, gamma_std: float):
        self.prob = prob
        self.gamma_std = gamma_std

    def draw(self, genparams: dict = {}):
        if not (np.random.rand() < self.prob or "gamma" in genparams.keys()):
            return None
        return np.exp(self.gamma_std * np.random.randn(1)[0]) if "gamma" not in genparams.keys() else genparams["gamma"]

    def __call__(self, output, device, genparams: dict = {}):
        from ...engine import SamplePlan

        gamma = self.draw(genparams)
        if gamma is None:
            return output, {"gamma": None}
        eng = _engine(device, output.shape)
        src = output.to(eng.device, torch.float32).contiguous()
        dst = torch.empty_like(src)
        eng.warp([SamplePlan(gamma=gamma)], [src.view(-1)], None, [dst.view(-1)], None)
        return dst, {"gamma": gamma}
