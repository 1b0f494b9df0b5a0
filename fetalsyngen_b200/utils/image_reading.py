"""Volume reader with the interface of the reference's ``SimpleITKReader``
(``fetalsyngen/utils/image_reading.py:8-55``): ``reader(path) -> tensor[x, y, z]``.
Backed by the dependency-free NIfTI codec (SimpleITK / monai are not required)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from .nifti import read_nifti, to_ras


class SimpleITKReader:
    """Volumes come back in RAS storage order like the reference's (SimpleITK read + monai ``Orientation("RAS")``,
    image_reading.py:32-55 with datasets.py:284-286): the axis-0 flip of the deformation is then a left-right flip
    whatever orientation the file was exported in."""

    def __call__(self, img_path: str | Path, as_meta: bool = True) -> torch.Tensor:
        arr, aff = to_ras(*read_nifti(img_path, with_affine=True))
        t = torch.from_numpy(np.ascontiguousarray(arr))
        # monai's MetaTensor is optional: attach the affine as a plain attribute
        try:
            t.affine = torch.from_numpy(aff)
        except Exception:
            pass
        return t
