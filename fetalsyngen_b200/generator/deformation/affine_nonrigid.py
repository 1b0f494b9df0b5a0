"""Affine + non-rigid spatial deformation.  Same class name, constructor and method contract as
the reference's ``SpatialDeformation``
(``fetalsyngen/generator/deformation/affine_nonrigid.py:12-366``).  The reference materialises
six coordinate grids, a 3-channel displacement field and int64 gather indices (~1.5 GB at
256^3); here the host only draws the parameters and ``fsg_warp`` does everything else in one
launch (control-grid up-sampling, affine, clamp, trilinear image + nearest segmentation)."""
from __future__ import annotations

from typing import Iterable

import numpy as np
import torch

from ...tables import make_affine_matrix


class SpatialDeformation:
    def __init__(
        self,
        max_rotation: float,
        max_shear: float,
        max_scaling: float,
        size: Iterable[int],
        prob: float,
        nonlinear_transform: bool,
        nonlin_scale_min: float,
        nonlin_scale_max: float,
        nonlin_std_max: float,
        flip_prb: float,
        device: str,
    ):
        self.size = size
        self.prob = prob
        self.flip_prb = flip_prb
        self.max_rotation = max_rotation
        self.max_shear = max_shear
        self.max_scaling = max_scaling
        self.nonlinear_transform = nonlinear_transform
        self.nonlin_scale_min = nonlin_scale_min
        self.nonlin_scale_max = nonlin_scale_max
        self.nonlin_std_max = nonlin_std_max
        self.device = device

    # ------------------------------------------------------------------ host draws
    def _shape_constants(self, shape3):
        """(centre2 f32, max_shift f32, center f32, shape as float64) of an input shape, built once."""
        cache = self.__dict__.setdefault("_shape_cache", {})
        c = cache.get(shape3)
        if c is None:
            shp = np.array(shape3)
            max_shift = ((shp - np.array(self.size)).astype(np.float32)) / 2
            max_shift[max_shift < 0] = 0
            c = (((shp - 1) / 2).astype(np.float32), max_shift, ((np.array(self.size) - 1) / 2).astype(np.float32), np.array(shape3))
            cache[shape3] = c
        return c

    def draw(self, image_shape, genparams: dict = {}, inject: dict | None = None, random_shift: bool = True, device_grids: bool = False):
        """Draw gate, flip, affine and control grid in the reference's RNG order
        (affine_nonrigid.py:140-145, 249-324).  Returns (plan_fields dict, deform_params dict)."""
        inject = inject or {}
        if not (np.random.rand() < self.prob or len(genparams.keys()) > 0):
            return {"deform": False, "flip": False}, {"affine": None, "non_rigid": None, "flip": False}
        flip = np.random.rand() < self.flip_prb if "flip" not in genparams.keys() else genparams["flip"]
        aff = genparams.get("affine", {})
        rotations = ((2 * self.max_rotation * np.random.rand(3) - self.max_rotation) / 180.0 * np.pi) if "rotations" not in aff.keys() else aff["rotations"]
        shears = (2 * self.max_shear * np.random.rand(3) - self.max_shear) if "shears" not in aff.keys() else aff["shears"]
        scalings = (1 + (2 * self.max_scaling * np.random.rand(3) - self.max_scaling)) if "scalings" not in aff.keys() else aff["scalings"]
        A = make_affine_matrix(rotations, shears, scalings).astype(np.float32)
        centre2, max_shift, center, shp_f = self._shape_constants(tuple(int(v) for v in image_shape[0:3]))
        if random_shift:
            u = np.asarray(inject["c2_u"], dtype=np.float64) if "c2_u" in inject else torch.rand(3, dtype=float).numpy()
            c2 = centre2 + (2 * (max_shift * u) - max_shift)  # float32 + float64 -> float64
        else:
            c2 = centre2.astype(np.float64)
        fields = {
            "deform": True,
            "flip": bool(flip),
            "A": A,
            "c2": np.asarray(c2, dtype=np.float64),
            "center": center,
            "fsmall": None,
        }
        non_rigid_params = {}
        if self.nonlinear_transform:
            nr = genparams.get("non_rigid", {})
            nonlin_scale = (self.nonlin_scale_min + np.random.rand(1) * (self.nonlin_scale_max - self.nonlin_scale_min)) if "nonlin_scale" not in nr.keys() else nr["nonlin_scale"]
            size_F_small = np.round(nonlin_scale * shp_f).astype(int).tolist() if "size_F_small" not in nr.keys() else nr["size_F_small"]
            nonlin_std = self.nonlin_std_max * np.random.rand() if "nonlin_std" not in nr.keys() else nr["nonlin_std"]
            if device_grids and "Fsmall_n" not in inject:
                # batched generator: the control grid is drawn on the device (fsg_draw_grids)
                fields["fsmall_dev"] = (tuple(int(v) for v in size_F_small), float(np.float32(nonlin_std)))
            else:
                if "Fsmall_n" in inject:
                    n = np.asarray(inject["Fsmall_n"], dtype=np.float32)
                else:
                    n = torch.randn([*size_F_small, 3], dtype=torch.float).numpy()
                fields["fsmall"] = (np.float32(nonlin_std) * n).astype(np.float32)
            non_rigid_params = {"nonlin_scale": nonlin_scale, "nonlin_std": nonlin_std, "size_F_small": size_F_small}
        params = {
            "affine": {"rotations": rotations, "shears": shears, "scalings": scalings},
            "non_rigid": non_rigid_params,
            "flip": flip,
        }
        return fields, params

    # ------------------------------------------------------------------ reference-compatible API
    def deform(self, image, segmentation, output, genparams: dict = {}, inject: dict | None = None):
        """Deform image / segmentation / output with one drawn transformation
        (affine_nonrigid.py:86-120).  Tensors stay on the device; dtypes are preserved."""
        from ...engine import SamplePlan, engine_for

        eng = engine_for(self.device, tuple(output.shape))
        fields, params = self.draw(output.shape, genparams, inject)
        plan = SamplePlan(**fields)
        seg_dtype = segmentation.dtype
        seg_u8 = eng.to_u8(segmentation.to(eng.device))
        src = output.to(eng.device, torch.float32).contiguous()
        dst = torch.empty_like(src)
        dseg = torch.empty_like(seg_u8)
        src2 = dst2 = None
        if image is not None:
            src2 = image.to(eng.device, torch.float32).contiguous()
            dst2 = torch.empty_like(src2)
        eng.warp([plan], [src.view(-1)], [seg_u8.view(-1)], [dst.view(-1)], [dseg.view(-1)],
                 None if src2 is None else [src2.view(-1)], None if dst2 is None else [dst2.view(-1)], epilogue=False)
        return dst2, eng.from_u8(dseg, seg_dtype), dst, params
