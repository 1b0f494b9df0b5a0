"""TEST INFRASTRUCTURE — loads the *unmodified* reference from /root/reference.

The reference (Medical-Image-Analysis-Laboratory/fetalsyngen) imports packages that are not
in this image (monai, SimpleITK, hydra, scikit-image, nibabel).  This module registers
minimal stand-ins in ``sys.modules`` so that ``import fetalsyngen`` works from the read-only
tree, and offers helpers to build the reference generator from its own YAML config.

It is used only (a) by ``tests/golden/make_golden.py`` to produce the committed golden
vectors and (b) by the in-container test that pins ``oracle/np_oracle.py`` against the
reference.  ``/root/reference`` does not exist on the GPU box: nothing in the ``-m gpu``
tests, ``smoke()`` or ``bench.py`` imports this file.

Stand-ins (what the reference needs from each package on the generation path):
  monai.transforms.Transform            base class only (synthseg.py:13)
  monai.transforms.Orientation          identity: bundled volumes are RAS already (sform=+0.5*I)
  monai.transforms.ScaleIntensity       (x-min)/(max-min)*(maxv-minv)+minv  (datasets.py:40,311)
  monai.transforms.Compose / Spacing    imported, never called on this path
  monai.data.MetaTensor                 torch.Tensor subclass (image_reading.py:32 annotation)
  SimpleITK.ReadImage/GetArrayFromImage gzip+struct NIfTI-1 reader, array returned (z,y,x)
  hydra.utils.instantiate               recursive ``_target_`` builder over yaml.safe_load
  skimage.morphology.ball               x^2+y^2+z^2 <= r^2 on a (2r+1)^3 grid
  nibabel                               empty module
"""
from __future__ import annotations

import gzip
import importlib
import os
import struct
import sys
import types
from pathlib import Path

import numpy as np
import torch
import yaml

REF_ROOT = Path(os.environ.get("FSG_REFERENCE_ROOT", "/root/reference"))


def reference_available() -> bool:
    return (REF_ROOT / "fetalsyngen" / "generator" / "model.py").exists()


# --------------------------------------------------------------------------- NIfTI-1
def read_nifti(path) -> tuple[np.ndarray, np.ndarray]:
    """Return (array in (x,y,z) order, 4x4 sform affine)."""
    raw = gzip.open(str(path), "rb").read() if str(path).endswith(".gz") else Path(path).read_bytes()
    dim = struct.unpack("<8h", raw[40:56])
    dtype_code = struct.unpack("<h", raw[70:72])[0]
    vox_offset = int(struct.unpack("<f", raw[108:112])[0])
    srow = np.array(struct.unpack("<12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
    np_dtype = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16}[dtype_code]
    n = int(np.prod(dim[1 : 1 + dim[0]]))
    data = np.frombuffer(raw, dtype=np_dtype, count=n, offset=vox_offset)
    arr = data.reshape(dim[1:4], order="F")
    aff = np.eye(4)
    aff[:3] = srow
    return arr, aff


class _SitkImage:
    def __init__(self, arr_xyz, aff):
        self.arr_xyz = arr_xyz
        self.aff = aff

    def GetDepth(self):
        return self.arr_xyz.shape[2]

    def TransformContinuousIndexToPhysicalPoint(self, p):
        # SimpleITK physical space is LPS: flip x,y of the RAS sform
        v = self.aff @ np.array([p[0], p[1], p[2], 1.0])
        return (-v[0], -v[1], v[2])


def _install_stubs():
    if "monai" in sys.modules and getattr(sys.modules["monai"], "_fsg_stub", False):
        return

    # ---- monai
    monai = types.ModuleType("monai")
    monai._fsg_stub = True
    mt = types.ModuleType("monai.transforms")
    md = types.ModuleType("monai.data")

    class Transform:
        pass

    class Orientation:
        def __init__(self, axcodes="RAS", **kw):
            self.axcodes = axcodes

        def __call__(self, x):
            return x

    class ScaleIntensity:
        def __init__(self, minv=0.0, maxv=1.0, **kw):
            self.minv, self.maxv = minv, maxv

        def __call__(self, img):
            mina, maxa = img.min(), img.max()
            if mina == maxa:
                return img * self.minv if self.minv is not None else img
            norm = (img - mina) / (maxa - mina)
            return norm * (self.maxv - self.minv) + self.minv

    class Compose:
        def __init__(self, transforms=None, **kw):
            self.transforms = transforms or []

        def __call__(self, x):
            for t in self.transforms:
                x = t(x)
            return x

    class Spacing:
        def __init__(self, *a, **kw):
            pass

    class MetaTensor(torch.Tensor):
        def __new__(cls, x=None, affine=None, **kw):
            t = torch.as_tensor(x)
            return t  # plain tensor: metadata is never read on the generation path

    mt.Transform, mt.Orientation, mt.ScaleIntensity = Transform, Orientation, ScaleIntensity
    mt.Compose, mt.Spacing = Compose, Spacing
    md.MetaTensor = MetaTensor
    monai.transforms, monai.data = mt, md
    sys.modules.update({"monai": monai, "monai.transforms": mt, "monai.data": md})

    # ---- SimpleITK
    sitk = types.ModuleType("SimpleITK")

    def ReadImage(path):
        arr, aff = read_nifti(path)
        return _SitkImage(arr, aff)

    def GetArrayFromImage(img):
        return np.ascontiguousarray(np.transpose(img.arr_xyz, (2, 1, 0)))

    sitk.ReadImage, sitk.GetArrayFromImage = ReadImage, GetArrayFromImage
    sys.modules["SimpleITK"] = sitk

    # ---- hydra
    hydra = types.ModuleType("hydra")
    hu = types.ModuleType("hydra.utils")
    hu.instantiate = instantiate
    hydra.utils = hu
    hydra.main = lambda *a, **k: (lambda f: f)
    sys.modules.update({"hydra": hydra, "hydra.utils": hu})

    # ---- skimage
    sk = types.ModuleType("skimage")
    skm = types.ModuleType("skimage.morphology")

    def ball(radius, dtype=np.uint8):
        n = 2 * radius + 1
        z, y, x = np.mgrid[-radius : radius : n * 1j, -radius : radius : n * 1j, -radius : radius : n * 1j]
        return (x * x + y * y + z * z <= radius * radius).astype(dtype)

    skm.ball = ball
    sk.morphology = skm
    sys.modules.update({"skimage": sk, "skimage.morphology": skm})

    sys.modules.setdefault("nibabel", types.ModuleType("nibabel"))


# --------------------------------------------------------------------------- _target_ builder
def _resolve(node, root, path):
    """Resolve ``${..key}`` relative interpolations (the only form the configs use)."""
    if isinstance(node, str) and node.startswith("${") and node.endswith("}"):
        ref = node[2:-1]
        up = len(ref) - len(ref.lstrip("."))
        key = ref.lstrip(".")
        base = path[: len(path) - (up - 1)] if up > 0 else []
        cur = root
        for k in base:
            cur = cur[k]
        return cur[key]
    return node


def instantiate(cfg, _root=None, _path=None, **overrides):
    """Minimal stand-in for ``hydra.utils.instantiate`` (recursive ``_target_`` construction)."""
    root = cfg if _root is None else _root
    path = [] if _path is None else _path
    if isinstance(cfg, dict):
        built = {}
        for k, v in cfg.items():
            if k == "_target_":
                continue
            v = _resolve(v, root, path)
            built[k] = instantiate(v, root, path + [k]) if isinstance(v, (dict, list)) else v
        built.update(overrides)
        if "_target_" in cfg:
            mod, _, name = cfg["_target_"].rpartition(".")
            return getattr(importlib.import_module(mod), name)(**built)
        return built
    if isinstance(cfg, list):
        return [instantiate(_resolve(v, root, path), root, path + [i]) if isinstance(v, (dict, list)) else _resolve(v, root, path) for i, v in enumerate(cfg)]
    return cfg


def load_reference():
    """Import the reference package from the read-only tree; returns the ``fetalsyngen`` module."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install_stubs()
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    if str(REF_ROOT) not in sys.path:
        sys.path.insert(0, str(REF_ROOT))
    return importlib.import_module("fetalsyngen")


def reference_generator_config(device="cpu", shape=(256, 256, 256), resolution=(0.5, 0.5, 0.5), artifacts=False) -> dict:
    cfg = yaml.safe_load((REF_ROOT / "configs/dataset/generator/default.yaml").read_text())
    cfg["device"] = device
    cfg["shape"] = list(shape)
    cfg["resolution"] = list(resolution)
    if not artifacts:
        for k in ("blur_cortex", "struct_noise", "simulate_motion", "boundaries"):
            cfg.pop(k, None)
    return cfg


def seed_paths(subject="sub-sta30") -> dict:
    base = REF_ROOT / "data/derivatives/seeds"
    out = {}
    for n in range(1, 7):
        out[n] = {m: next((base / f"subclasses_{n}" / subject / "anat").glob(f"*_mlabel_{m}.nii.gz")) for m in range(1, 5)}
    return out


def seg_path(subject="sub-sta30") -> Path:
    return next((REF_ROOT / "data" / subject / "anat").glob("*_dseg.nii.gz"))
