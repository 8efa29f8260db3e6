"""In-tree build of libloupiote_b200.so (sm_100a only) with a plain nvcc command line.

The built library lives at loupiote_b200/_lib/libloupiote_b200.so so that it travels with
the source snapshot to the GPU box.  `python -m loupiote_b200._build` rebuilds it.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "loupiote_b200"
CSRC = PKG / "csrc"
LIB_DIR = PKG / "_lib"
LIB_PATH = LIB_DIR / "libloupiote_b200.so"

SOURCES = [
    CSRC / "host" / "bvh_build.cpp",
    CSRC / "host" / "scene.cpp",
    CSRC / "host" / "gltf.cpp",
    CSRC / "host" / "api_scene.cpp",
    CSRC / "cuda" / "api_render.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # FMA contraction is explicit in the kernels (arithmetic contract)
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3,-ffp-contract=off",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = sorted(list(CSRC.rglob("*.cu")) + list(CSRC.rglob("*.cuh")) + list(CSRC.rglob("*.cpp"))
                   + list(CSRC.rglob("*.hpp")) + [ROOT / "include" / "loupiote.h"])
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the library if sources changed since the last build; returns its path."""
    LIB_DIR.mkdir(exist_ok=True)
    stamp = LIB_DIR / "build.stamp"
    fp = _fingerprint()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, f"-I{ROOT / 'include'}", "-o", str(LIB_PATH)]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [str(s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libloupiote_b200.so")
    if verbose:
        sys.stderr.write(proc.stdout + proc.stderr)
    stamp.write_text(fp)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
