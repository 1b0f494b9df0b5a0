"""Volume reader with the interface of the reference's ``SimpleITKReader``
(``fetalsyngen/utils/image_reading.py:8-55``): ``reader(path) -> tensor[x, y, z]``.
Backed by the dependency-free NIfTI codec (SimpleITK / monai are not required)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from .nifti import read_nifti


class SimpleITKReader:
    def __call__(self, img_path: str | Path, as_meta: bool = True) -> torch.Tensor:
        arr, aff = read_nifti(img_path, with_affine=True)
        t = torch.from_numpy(np.ascontiguousarray(arr))
        # monai's MetaTensor is optional: attach the affine as a plain attribute
        try:
            t.affine = torch.from_numpy(aff)
        except Exception:
            pass
        return t
