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
OUT = ROOT / "_ref"


def build(verbose: bool = False):
    if not (REF / "slice_acq_cuda.cpp").exists():
        print("reference sources not found; nothing built")
        return None
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    OUT.mkdir(exist_ok=True)
    return load("slice_acq_cuda", [str(REF / "slice_acq_cuda.cpp"), str(REF / "slice_acq_cuda_kernel.cu")], build_directory=str(OUT), verbose=verbose, is_python_module=True)


def load_built():
    """Import the pre-built module from oracle/_ref (no compiler needed)."""
    so = OUT / "slice_acq_cuda.so"
    if not so.exists():
        return None
    import importlib.util

    import torch  # noqa: F401  (the extension links against libtorch)

    spec = importlib.util.spec_from_file_location("slice_acq_cuda", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    m = build(verbose="-v" in sys.argv)
    print("built:", m, sorted(p.name for p in OUT.iterdir()))
