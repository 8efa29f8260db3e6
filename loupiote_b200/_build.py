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
    CSRC / "host" / "image_decode.cpp",
    CSRC / "host" / "textures.cpp",
    CSRC / "host" / "api_scene.cpp",
    CSRC / "cuda" / "api_render.cu",
    CSRC / "cuda" / "api_multi.cu",
    CSRC / "cuda" / "lbvh_build.cu",
]
# translation units whose arithmetic never decides a hit: FMA contraction on
SOURCES_FMAD = [
    CSRC / "cuda" / "shade_kernel.cu",
    CSRC / "cuda" / "svgf_kernels.cu",
]

COMMON_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3,-ffp-contract=off",
]
# FMA contraction is explicit in the kernels (arithmetic contract)
NVCC_FLAGS = COMMON_FLAGS + ["-fmad=false", "-prec-div=true", "-prec-sqrt=true"]
NVCC_FLAGS_FMAD = COMMON_FLAGS + ["-fmad=true", "-prec-div=true", "-prec-sqrt=true"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = sorted([ROOT / "tools" / "cli" / "lp_render.cpp"] + list(CSRC.rglob("*.cu")) + list(CSRC.rglob("*.cuh")) + list(CSRC.rglob("*.cpp")) + list(CSRC.rglob("*.h"))
                   + list(CSRC.rglob("*.hpp")) + [ROOT / "include" / "loupiote.h"])
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS + NVCC_FLAGS_FMAD).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, variant: str = "",
          defines: tuple = ()) -> Path:
    """Compile the library if sources changed since the last build; returns its path.

    `variant` / `defines` build a tuning copy libloupiote_b200.<variant>.so with extra -D
    flags (A/B measurements on the GPU box: LP_LIB_VARIANT=<variant> selects it at load)."""
    LIB_DIR.mkdir(exist_ok=True)
    lib_path = LIB_DIR / (f"libloupiote_b200.{variant}.so" if variant else LIB_PATH.name)
    stamp = LIB_DIR / (f"build.{variant}.stamp" if variant else "build.stamp")
    fp = _fingerprint() + " " + " ".join(defines)
    if not force and lib_path.exists() and stamp.exists() and stamp.read_text() == fp:
        return lib_path
    obj_dir = LIB_DIR / (f"obj.{variant}" if variant else "obj")
    obj_dir.mkdir(exist_ok=True)
    jobs = [(src, NVCC_FLAGS) for src in SOURCES] + [(src, NVCC_FLAGS_FMAD) for src in SOURCES_FMAD]
    procs, objs = [], []
    for src, flags in jobs:  # one nvcc per translation unit, all in parallel
        obj = obj_dir / (src.stem + ".o")
        objs.append(obj)
        cmd = [_nvcc(), *flags, *[f"-D{d}" for d in defines], f"-I{ROOT / 'include'}", "-c",
               str(src), "-o", str(obj)]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                      text=True))
    failed = False
    for proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0 or verbose:
            sys.stderr.write(out)
        failed |= proc.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libloupiote_b200.so")
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(lib_path),
            *[str(o) for o in objs], "-lnccl"]  # lp_multi_*: NCCL over NVLink
    proc = subprocess.run(link, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed linking libloupiote_b200.so")
    if not variant:
        build_cli(lib_path)
    stamp.write_text(fp)
    return lib_path


CLI_SRC = ROOT / "tools" / "cli" / "lp_render.cpp"
CLI_PATH = LIB_DIR / "lp_render"


def build_cli(lib_path: Path = LIB_PATH) -> Path:
    """Headless renderer over the C ABI only (tools/cli/lp_render.cpp), linked against the
    in-tree library with an $ORIGIN rpath."""
    cmd = ["g++", "-O2", "-std=c++17", f"-I{ROOT / 'include'}", str(CLI_SRC), "-o", str(CLI_PATH),
           f"-L{lib_path.parent}", "-l:" + lib_path.name, "-Wl,-rpath,$ORIGIN"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("g++ failed building lp_render")
    return CLI_PATH


if __name__ == "__main__":
    # python -m loupiote_b200._build [--force] [-v] [--variant NAME -DX=1 -DY ...]
    _variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    _defs = tuple(a[2:] for a in sys.argv if a.startswith("-D"))
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=_variant,
                defines=_defs))
