#!/usr/bin/env python
"""Mutation fuzzing of the ingest path that takes untrusted bytes: lp_load_gltf
[ref loaders/gltf.rs:46-156] and the PNG / JPEG decoders behind it [ref gltf.rs:12-44].

Every mutated input must come back as LP_OK or an error status -- never a crash, never a
sanitizer report.  Seeds: the reference's one fixture (tests/golden/cornell-box.glb) and the
image fixtures of tests/golden/image_fixtures.npz; mutations: byte overwrites, bit flips,
truncation, extreme 16/32-bit values, with PNG chunk CRCs recomputed most of the time so that
the mutation reaches the decoder proper.

    python tools/fuzz_ingest.py --n 3000 --seed 1            # the product library
    python tools/fuzz_ingest.py --n 3000 --seed 1 --asan     # host sources rebuilt with
                                                             # -fsanitize=address,undefined
    python tools/fuzz_ingest.py --n 3000 --structured [--asan]   # well-formed JSON, mutated
                                                             # meaning (offsets, counts, types)
(--asan re-executes itself with libasan preloaded; the sanitized library holds the host
sources only, which is all the ingest path needs.)
"""
import argparse
import copy
import ctypes as C
import json
import os
import random
import struct
import subprocess
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ASAN_LIB = ROOT / "oracle" / "_build" / "libloupiote_host_asan.so"


def fix_png_crcs(b: bytes) -> bytes:
    if b[:8] != b"\x89PNG\r\n\x1a\n":
        return b
    out, p = bytearray(b[:8]), 8
    while p + 8 <= len(b):
        ln = struct.unpack(">I", b[p:p + 4])[0]
        if p + 12 + ln > len(b):
            break
        typ, body = b[p + 4:p + 8], b[p + 8:p + 8 + ln]
        out += b[p:p + 8] + body + struct.pack(">I", zlib.crc32(typ + body) & 0xFFFFFFFF)
        p += 12 + ln
    return bytes(out + b[p:])


def mutate(rng: random.Random, data: bytes) -> bytes:
    b, mode = bytearray(data), rng.randrange(5)
    if mode == 0:
        for _ in range(rng.randrange(1, 6)):
            b[rng.randrange(len(b))] = rng.randrange(256)
    elif mode == 1:
        b = b[:rng.randrange(len(b))]
    elif mode == 2:
        p = rng.randrange(len(b) - 2)
        b[p:p + 2] = bytes([rng.choice([0, 255, 0x7F, 0x80]), rng.choice([0, 255, 1])])
    elif mode == 3:
        for _ in range(rng.randrange(1, 4)):
            b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
    elif len(b) > 8:
        struct.pack_into("<I", b, rng.randrange(len(b) - 4),
                         rng.choice([0, 0xFFFFFFFF, 0x7FFFFFFF, len(b), len(b) + 1]))
    return bytes(b)


def split_glb(glb: bytes):
    jl = struct.unpack_from("<I", glb, 12)[0]
    doc = json.loads(glb[20:20 + jl])
    rest = glb[20 + jl:]
    blob = rest[8:8 + struct.unpack_from("<I", rest, 0)[0]] if len(rest) >= 8 else b""
    return doc, blob


def join_glb(doc, blob: bytes) -> bytes:
    js = json.dumps(doc).encode()
    js += b" " * (-len(js) % 4)
    total = 12 + 8 + len(js) + 8 + len(blob)
    return (b"glTF" + struct.pack("<II", 2, total) + struct.pack("<I", len(js)) + b"JSON" + js +
            struct.pack("<I", len(blob)) + b"BIN\x00" + blob)


def mutate_json(rng: random.Random, node):
    """Structure-aware mutation: walks to a random place of the glTF document and replaces a
    value with an extreme of the same or another type, drops a key, or duplicates an item --
    the JSON stays well formed, its meaning does not (offsets past the buffer, negative or
    huge counts and indices, wrong component types, cyclic / missing references)."""
    path = []
    cur = node
    while isinstance(cur, (dict, list)) and cur and (not path or rng.random() < 0.85):
        key = rng.choice(list(cur.keys())) if isinstance(cur, dict) else rng.randrange(len(cur))
        if not path and rng.random() < 0.9:  # mostly where the loader follows references
            hot = [k for k in ("accessors", "bufferViews", "meshes", "nodes", "materials",
                               "textures", "images", "buffers") if k in cur]
            key = rng.choice(hot) if hot else key
        path.append((cur, key))
        cur = cur[key]
    if not path:
        return
    parent, key = path[-1]
    extremes = [-1, 0, 1, 2, 255, 65535, 65536, 2 ** 31 - 1, 2 ** 32 - 1, 2 ** 32, 10 ** 12, -2 ** 31,
                0.5, -0.0, 1e38, 1e308, None, "", "VEC3", "MAT4", "SCALAR", [], {}, True,
                5120, 5121, 5123, 5125, 5126, 34962]
    mode = rng.randrange(4)
    if mode == 0 or not isinstance(parent, (dict, list)):
        parent[key] = rng.choice(extremes)
    elif mode == 1 and isinstance(parent, dict):
        del parent[key]
    elif mode == 2 and isinstance(parent, list):
        parent.append(parent[key])
    elif isinstance(parent[key], (int, float)) and not isinstance(parent[key], bool):
        parent[key] = parent[key] + rng.choice([-1, 1, 3, -7, 1000, 0.25])
    else:
        parent[key] = rng.choice(extremes)


def run_structured(lib: C.CDLL, n: int, seed: int) -> dict:
    """Well-formed GLBs whose JSON has been mutated structurally (mutate_json), from two seeds:
    the reference's cornell box and a textured two-material GLB with embedded PNG images."""
    sys.path.insert(0, str(ROOT / "tests"))
    from _glb import textured_quad_glb
    z = np.load(ROOT / "tests" / "golden" / "image_fixtures.npz")
    png = bytes(z["file_png_rgb8"].tobytes())
    seeds = [(ROOT / "tests" / "golden" / "cornell-box.glb").read_bytes(),
             textured_quad_glb([png, png], [0, 1], [{"base": 0, "mr": 1}, {"base": 1}])]
    parts = [split_glb(g) for g in seeds]
    lib.lp_load_gltf.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
    lib.lp_scene_get_array.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    rng = random.Random(seed)
    counts = {"ok": 0, "rejected": 0}
    for _ in range(n):
        doc, blob = parts[rng.randrange(len(parts))]
        doc = copy.deepcopy(doc)
        for _ in range(rng.randrange(1, 4)):
            mutate_json(rng, doc)
        if rng.random() < 0.1:
            blob = blob[:rng.randrange(len(blob) + 1)]
        b = join_glb(doc, blob)
        h = C.c_void_p()
        assert lib.lp_scene_create(C.byref(h)) == 0
        st = lib.lp_load_gltf(b, len(b), h)
        if st == 0:  # what was accepted must also survive the derived builds (TLAS, layouts)
            ptr, cnt, es = C.c_void_p(), C.c_size_t(), C.c_size_t()
            lib.lp_scene_get_array(h, 12, C.byref(ptr), C.byref(cnt), C.byref(es))
        counts["ok" if st == 0 else "rejected"] += 1
        lib.lp_scene_destroy(h)
    return counts


def run_api(lib: C.CDLL, n: int, seed: int) -> dict:
    """The scene-building entry points with extreme but well-typed arguments: vertex positions
    and instance matrices drawn from {0, -0, denormals, +-1e19, +-3e38, NaN, inf, ...}, empty
    and out-of-range index lists; every accepted scene then goes through the TLAS build and
    the GPU re-layouts (lp_scene_get_array of the derived arrays)."""
    void_p, size_t, u32p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)
    lib.lp_scene_add_bvh.argtypes = [void_p, void_p, size_t, void_p, size_t, void_p, size_t, size_t, u32p]
    lib.lp_scene_add_bvh_indexed.argtypes = [void_p, void_p, size_t, void_p, size_t, void_p, size_t,
                                             size_t, void_p, size_t, u32p]
    lib.lp_scene_add_instance.argtypes = [void_p, C.c_uint32, void_p, C.c_uint32]
    lib.lp_scene_get_array.argtypes = [void_p, C.c_int, C.POINTER(void_p), C.POINTER(size_t),
                                       C.POINTER(size_t)]
    rng = np.random.default_rng(seed)
    vals = np.array([0, -0.0, 1e-45, -1e-38, 1, -1, 1e19, -1e19, 3e38, -3e38, 65504, 1e-3, 7.5],
                    np.float32)
    counts = {"ok": 0, "rejected": 0}
    for _ in range(n):
        h = void_p()
        assert lib.lp_scene_create(C.byref(h)) == 0
        for _b in range(int(rng.integers(1, 4))):
            nv = int(rng.integers(0, 40)) * 3
            mode = int(rng.integers(0, 4))
            if mode == 0:
                pos = rng.normal(size=(nv, 3)).astype(np.float32)
            elif mode == 1:
                pos = rng.choice(vals, size=(nv, 3)).astype(np.float32)
            elif mode == 2:
                with np.errstate(over="ignore"):
                    pos = (rng.normal(size=(nv, 3)) * 10.0 ** int(rng.integers(-30, 38))).astype(np.float32)
            else:
                pos = rng.normal(size=(nv, 3)).astype(np.float32)
                if nv:
                    pos[rng.integers(0, nv), rng.integers(0, 3)] = rng.choice([np.nan, np.inf, -np.inf])
            idx = C.c_uint32()
            if rng.random() < 0.5 and nv:
                hi = nv + (2 if rng.random() < 0.2 else 0)
                ind = rng.integers(0, hi, size=int(rng.integers(0, 30)) * 3).astype(np.uint32)
                st = lib.lp_scene_add_bvh_indexed(h, pos.ctypes.data, 12, None, 0, None, 0, nv,
                                                  ind.ctypes.data, len(ind), C.byref(idx))
            else:
                st = lib.lp_scene_add_bvh(h, pos.ctypes.data if nv else None, 12, None, 0, None, 0,
                                          nv, C.byref(idx))
            if st != 0:
                continue
            for _k in range(int(rng.integers(0, 4))):
                m = np.eye(4, dtype=np.float32)
                mm = int(rng.integers(0, 4))
                if mm == 0:
                    m[:3, :3] = rng.normal(size=(3, 3))
                elif mm == 1:
                    m = rng.choice(vals, size=(4, 4)).astype(np.float32)
                elif mm == 2:
                    m[:3, :3] *= np.float32(10.0 ** int(rng.integers(-30, 30)))
                else:
                    m[rng.integers(0, 4), rng.integers(0, 4)] = rng.choice([np.nan, np.inf])
                m = np.ascontiguousarray(m)
                lib.lp_scene_add_instance(h, idx.value, m.ctypes.data, 0)
        ptr, cnt, es = void_p(), size_t(), size_t()
        st = 0
        for which in (12, 15, 10, 9, 11):
            st |= lib.lp_scene_get_array(h, which, C.byref(ptr), C.byref(cnt), C.byref(es))
        counts["ok" if st == 0 else "rejected"] += 1
        lib.lp_scene_destroy(h)
    return counts


def run(lib: C.CDLL, n: int, seed: int) -> dict:
    z = np.load(ROOT / "tests" / "golden" / "image_fixtures.npz")
    images = [bytes(z[k].tobytes()) for k in sorted(z.keys()) if k.startswith("file_")]
    glb = (ROOT / "tests" / "golden" / "cornell-box.glb").read_bytes()
    lib.lp_scene_push_encoded_image.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t,
                                                C.POINTER(C.c_uint32)]
    lib.lp_load_gltf.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
    rng = random.Random(seed)
    counts = {"ok": 0, "rejected": 0}
    for _ in range(n):
        h = C.c_void_p()
        assert lib.lp_scene_create(C.byref(h)) == 0
        if rng.random() < 0.6:
            b = mutate(rng, rng.choice(images))
            if rng.random() < 0.7:
                b = fix_png_crcs(b)
            idx = C.c_uint32()
            st = lib.lp_scene_push_encoded_image(h, b, len(b), C.byref(idx))
        else:
            b = mutate(rng, glb)
            st = lib.lp_load_gltf(b, len(b), h)
        counts["ok" if st == 0 else "rejected"] += 1
        lib.lp_scene_destroy(h)
    return counts


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--asan", action="store_true")
    ap.add_argument("--api", action="store_true",
                    help="scene-building entry points with extreme float arguments")
    ap.add_argument("--structured", action="store_true",
                    help="structure-aware mutations of the glTF JSON instead of byte mutations")
    args = ap.parse_args()
    if args.asan and os.environ.get("LP_FUZZ_CHILD") != "1":
        ASAN_LIB.parent.mkdir(exist_ok=True)
        srcs = sorted(str(p) for p in (ROOT / "loupiote_b200" / "csrc" / "host").glob("*.cpp"))
        subprocess.run(["g++", "-O1", "-g", "-fsanitize=address,undefined,float-cast-overflow", "-fno-sanitize-recover=undefined,float-cast-overflow", "-fno-omit-frame-pointer",
                        "-std=c++17", "-shared", "-fPIC", f"-I{ROOT / 'include'}", *srcs, "-o",
                        str(ASAN_LIB)], check=True)
        asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True,
                              text=True, check=True).stdout.strip()
        cxx = subprocess.run(["gcc", "-print-file-name=libstdc++.so.6"], capture_output=True,
                             text=True, check=True).stdout.strip()
        env = dict(os.environ, LP_FUZZ_CHILD="1", LD_PRELOAD=f"{asan} {cxx}",
                   ASAN_OPTIONS="detect_leaks=0:abort_on_error=1")
        sys.exit(subprocess.run([sys.executable, __file__, *sys.argv[1:]], env=env).returncode)
    if args.asan:
        lib = C.CDLL(str(ASAN_LIB))
    else:
        sys.path.insert(0, str(ROOT))
        from loupiote_b200 import _ffi
        lib = _ffi.lib()
    counts = (run_api if args.api else run_structured if args.structured else run)(
        lib, args.n, args.seed)
    print({"n": args.n, "seed": args.seed, "sanitized": bool(args.asan),
           "structured": bool(args.structured), "api": bool(args.api), **counts})


if __name__ == "__main__":
    main()
