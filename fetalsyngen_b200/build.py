"""Build libfsg.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).
Every .cu is compiled to its own object (in parallel, re-used while it is newer than its sources and the
shared headers), then linked."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libfsg.so"
OBJ = ROOT / "build" / "obj"
SOURCES = ["core.cu", "gmm.cu", "warp.cu", "warp_tile.cu", "warp_pipe.cu", "blur.cu", "resample.cu", "sepconv.cu", "zoom.cu", "artifacts.cu", "motion.cu", "motion_ex.cu", "seeds.cu", "plan.cu", "texvol.cu", "step.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # parity: every float32 mul/add rounded separately, like torch eager
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libfsg.so cannot be built")


def _headers():
    return sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "fsg.h"]


def _sources():
    return [s for s in SOURCES if (CSRC / s).exists()]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(d.stat().st_mtime > t for d in [CSRC / s for s in _sources()] + _headers())


def build(force: bool = False, verbose: bool = False, out: Path | None = None, defines: tuple = ()) -> Path:
    """``out`` / ``defines``: variant builds for A/B kernel timing (``FSG_LIB`` selects one at load)."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    tag = "".join(c if c.isalnum() else "_" for c in "_".join(defines)) if defines else "default"
    objdir = OBJ / tag
    objdir.mkdir(parents=True, exist_ok=True)
    hdr_t = max(h.stat().st_mtime for h in _headers())
    logs = []

    def compile_one(src: str):
        o = objdir / (src + ".o")
        s = CSRC / src
        if not force and o.exists() and o.stat().st_mtime > max(s.stat().st_mtime, hdr_t):
            return o
        cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas=-v"] if verbose else []), *[f"-D{d}" for d in defines], f"-I{ROOT / 'include'}", "-c", str(s), "-o", str(o)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        logs.append(res.stderr)
        return o

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", *[str(o) for o in objs], "-o", str(out or LIB)], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    if verbose:
        print("".join(logs))
    return out or LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
