"""Build libfsg.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libfsg.so"
SOURCES = ["core.cu", "gmm.cu", "warp.cu", "warp_tile.cu", "blur.cu", "resample.cu", "sepconv.cu", "zoom.cu", "artifacts.cu", "motion.cu", "seeds.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # parity: every float32 mul/add rounded separately, like torch eager
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libfsg.so cannot be built")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [CSRC / "common.cuh", CSRC / "warp_common.cuh", ROOT / "include" / "fsg.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: Path | None = None, defines: tuple = ()) -> Path:
    """``out`` / ``defines``: variant builds for A/B kernel timing (``FSG_LIB`` selects one at load)."""
    if out is None and not force and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], f"-I{ROOT / 'include'}", *[str(CSRC / s) for s in SOURCES], "-o", str(out or LIB)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out or LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
