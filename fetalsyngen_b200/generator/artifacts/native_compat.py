"""Drop-in replacements of the reference's two pybind11 CUDA modules, over libfsg's C-ABI.

``slice_acq_cuda``  (svort/slice_acquisition/slice_acq_cuda.cpp:61-79,105-124): ``forward``, ``adjoint_forward``
``transform_convert_cuda``  (svort/transform/transform_convert_cuda.cpp:27-51): ``axisangle2mat_forward``,
``mat2axisangle_forward``

Same argument lists and return conventions as the JIT-built modules (CUDA float32 tensors in, lists of freshly
allocated tensors out, "optional" tensors passed as 0-element tensors, ``RuntimeError``-style failure for non-CUDA /
non-contiguous inputs), so ``slice_acq.py:12-19`` / ``transform_convert.py`` can import them instead of calling
``torch.utils.cpp_extension.load``::

    from fetalsyngen_b200.generator.artifacts.native_compat import slice_acq_cuda, transform_convert_cuda

Every option of the modules is covered (masks, ``need_weight``, both ``interp_psf`` modes, ``equalize``); the two call
shapes the generator itself uses take the tuned kernels of ``csrc/motion.cu``.  The backward kernels (autograd of SVoRT
training) are not part of the generation path and are not provided.
"""
from __future__ import annotations

import numpy as np
import torch

from ... import _lib
from . import svort


def _check(t: torch.Tensor, name: str, dtype=torch.float32):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}")
    return t


def _mask(t: torch.Tensor, name: str):
    """0-element tensor = absent (slice_acq.py:37-40); bool / uint8 masks are read as bytes."""
    if t is None or t.numel() == 0:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous CUDA tensor")
    return t.view(torch.uint8) if t.dtype == torch.bool else _check(t, name, torch.uint8)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _taps(psf: torch.Tensor, extra: float = 0.0):
    taps, radius = svort.psf_taps(psf.detach().cpu().numpy())
    return torch.from_numpy(np.ascontiguousarray(taps, dtype=np.float32)).to(psf.device), float(radius) + extra


class _SliceAcq:
    """``slice_acq_cuda.forward / adjoint_forward``."""

    @staticmethod
    def forward(transforms, vol, vol_mask, slices_mask, psf, slice_shape, res_slice, need_weight, interp_psf):
        for t, nm in ((transforms, "transforms"), (vol, "vol"), (psf, "psf")):
            _check(t, nm)
        n, (h, w) = int(transforms.shape[0]), (int(slice_shape[0]), int(slice_shape[1]))
        D, H, W = (int(s) for s in vol.shape[-3:])
        vm, sm = _mask(vol_mask, "vol_mask"), _mask(slices_mask, "slices_mask")
        taps, radius = _taps(psf, 1.0 if interp_psf else 0.0)
        slices = torch.empty((n, 1, h, w), dtype=torch.float32, device=vol.device)
        weights = torch.empty_like(slices) if need_weight else None
        dp, hp, wp = (int(s) for s in psf.shape)
        if vm is None and sm is None and not need_weight and not interp_psf:  # the generator's call: tuned kernel
            _lib.call("fsg_slice_acq_forward", transforms.data_ptr(), vol.data_ptr(), taps.data_ptr(), int(taps.shape[0]), radius, slices.data_ptr(), n, h, w, D, H, W,
                      float(res_slice), _stream())
        else:
            _lib.call("fsg_slice_acq_forward_ex", transforms.data_ptr(), vol.data_ptr(), None if vm is None else vm.data_ptr(), None if sm is None else sm.data_ptr(),
                      psf.data_ptr(), dp, hp, wp, taps.data_ptr(), int(taps.shape[0]), radius, slices.data_ptr(), None if weights is None else weights.data_ptr(), n, h, w,
                      D, H, W, float(res_slice), int(bool(interp_psf)), _stream())
        return [slices, weights] if need_weight else [slices]

    @staticmethod
    def adjoint_forward(transforms, psf, slices, slices_mask, vol_mask, vol_shape, res_slice, interp_psf, equalize):
        for t, nm in ((transforms, "transforms"), (slices, "slices"), (psf, "psf")):
            _check(t, nm)
        n = int(transforms.shape[0])
        h, w = (int(s) for s in slices.shape[-2:])
        D, H, W = (int(s) for s in vol_shape)
        vm, sm = _mask(vol_mask, "vol_mask"), _mask(slices_mask, "slices_mask")
        taps, radius = _taps(psf, 1.0)
        vol = torch.empty((1, 1, D, H, W), dtype=torch.float32, device=slices.device)
        wgt = torch.empty_like(vol) if equalize else None
        dp, hp, wp = (int(s) for s in psf.shape)
        if vm is None and sm is None and interp_psf:  # the generator's call: warp-per-pixel kernel, interleaved accumulator
            acc = torch.empty((D * H * W, 2), dtype=torch.float32, device=slices.device)
            _lib.call("fsg_slice_acq_adjoint", transforms.data_ptr(), psf.data_ptr(), dp, hp, wp, taps.data_ptr(), int(taps.shape[0]), radius, slices.data_ptr(), None,
                      vol.data_ptr(), None if wgt is None else wgt.data_ptr(), acc.data_ptr(), n, h, w, D, H, W, float(res_slice), int(bool(equalize)), _stream())
        else:
            _lib.call("fsg_slice_acq_adjoint_ex", transforms.data_ptr(), psf.data_ptr(), dp, hp, wp, taps.data_ptr(), int(taps.shape[0]), radius, slices.data_ptr(),
                      None if sm is None else sm.data_ptr(), None if vm is None else vm.data_ptr(), vol.data_ptr(), None if wgt is None else wgt.data_ptr(), n, h, w, D, H, W,
                      float(res_slice), int(bool(interp_psf)), int(bool(equalize)), _stream())
        return [vol, wgt if wgt is not None else torch.Tensor()]


class _TransformConvert:
    """``transform_convert_cuda.axisangle2mat_forward / mat2axisangle_forward``."""

    @staticmethod
    def axisangle2mat_forward(axisangle):
        _check(axisangle, "axisangle")
        n = int(axisangle.shape[0])
        mat = torch.empty((n, 3, 4), dtype=torch.float32, device=axisangle.device)
        if n:
            _lib.call("fsg_axisangle2mat", axisangle.data_ptr(), mat.data_ptr(), n, _stream())
        return [mat]

    @staticmethod
    def mat2axisangle_forward(mat):
        _check(mat, "mat")
        n = int(mat.shape[0])
        ax = torch.empty((n, 6), dtype=torch.float32, device=mat.device)
        if n:
            _lib.call("fsg_mat2axisangle", mat.data_ptr(), ax.data_ptr(), n, _stream())
        return [ax]


slice_acq_cuda = _SliceAcq()
transform_convert_cuda = _TransformConvert()
