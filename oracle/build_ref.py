"""TEST INFRASTRUCTURE — builds the reference's own CUDA extension for the motion artifact
(``fetalsyngen/generator/artifacts/svort/slice_acquisition/slice_acq_cuda{.cpp,_kernel.cu}``)
from the sources where they lie under /root/reference into ``oracle/_ref/`` (git-ignored; the
built ``.so`` travels to the GPU box with the repo snapshot).  No reference source is copied.

    python oracle/build_ref.py

Used by ``tests/test_gpu_motion.py`` as the ground truth for ``fsg_slice_acq_forward`` /
``fsg_slice_acq_adjoint``; skipped where the ``.so`` is absent.  The reference builds the same
sources with ``torch.utils.cpp_extension.load`` at import time (slice_acq.py:12-19).
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
REF = Path(os.environ.get("FSG_REFERENCE_ROOT", "/root/reference")) / "fetalsyngen/generator/artifacts/svort/slice_acquisition"
REF_T = REF.parent / "transform"  # transform_convert_cuda{.cpp,_kernel.cu}: the reference's second pybind module
OUT = ROOT / "_ref"


def build(verbose: bool = False):
    if not (REF / "slice_acq_cuda.cpp").exists():
        print("reference sources not found; nothing built")
        return None
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    OUT.mkdir(exist_ok=True)
    mod = load("slice_acq_cuda", [str(REF / "slice_acq_cuda.cpp"), str(REF / "slice_acq_cuda_kernel.cu")], build_directory=str(OUT), verbose=verbose, is_python_module=True)
    if (REF_T / "transform_convert_cuda.cpp").exists():
        (OUT / "tc").mkdir(exist_ok=True)
        load("transform_convert_cuda", [str(REF_T / "transform_convert_cuda.cpp"), str(REF_T / "transform_convert_cuda_kernel.cu")], build_directory=str(OUT / "tc"), verbose=verbose,
             is_python_module=True)
    return mod


def load_built(name: str = "slice_acq_cuda"):
    """Import a pre-built module from oracle/_ref (no compiler needed): ``slice_acq_cuda`` or ``transform_convert_cuda``."""
    so = (OUT / "slice_acq_cuda.so") if name == "slice_acq_cuda" else (OUT / "tc" / f"{name}.so")
    if not so.exists():
        return None
    import importlib.util

    import torch  # noqa: F401  (the extension links against libtorch)

    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    m = build(verbose="-v" in sys.argv)
    print("built:", m, sorted(p.name for p in OUT.iterdir()))
