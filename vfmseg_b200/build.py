"""In-tree build of libvfmseg_b200.so (sm_100a only) with plain nvcc.

The shared library is git-ignored but travels with the gpurun snapshot; `build()` is what
`__graft_entry__.build()` calls. nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
# VFMSEG_B200_LIB: experiment knob (tools/): load another build of the same sources, e.g. a macro sweep variant
LIB_PATH = Path(os.environ["VFMSEG_B200_LIB"]) if os.environ.get("VFMSEG_B200_LIB") else LIB_DIR / "libvfmseg_b200.so"
SOURCES = [CSRC / "api.cu"]
HEADERS = sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "vfmseg_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libvfmseg_b200.so")


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the CUDA library if missing or older than its sources; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    LIB_DIR.mkdir(exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB_PATH), *map(str, SOURCES)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
