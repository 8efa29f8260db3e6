"""CPU checks of the GPU BVH builder (loupiote_b200/csrc/cuda/lbvh_core.h, SURVEY 8(f) row 4;
replaces on the device the host build behind BLASArray::add_bvh, gltf.rs:97-105).

tests/lbvh_emu.cpp runs the builder's per-thread bodies -- the very functions the CUDA
kernels call -- in a serial loop.  A tree is valid iff every slot box is exactly the bounds
of what hangs below it and every triangle (instance) hangs below the root exactly once:
any conservative traversal of such a tree reaches every triangle a ray can hit, so results
equal the canonical tree's (closest hit is order independent).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import subprocess
from pathlib import Path

import numpy as np
import pytest

import loupiote_b200 as lb
from loupiote_b200 import _ffi, scenes

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "lbvh_emu.cpp"
CORE = ROOT / "loupiote_b200" / "csrc" / "cuda" / "lbvh_core.h"
OUT = ROOT / "oracle" / "_build" / "liblbvh_emu.so"

LEAF = 0x80000000
NONE = 0x7FFFFFFF
NODE2 = np.dtype([("q", "<f4", 12), ("child", "<u4", 2), ("pad", "<u4", 2)])
NODE4 = np.dtype([("lo", "<f4", (3, 4)), ("hi", "<f4", (3, 4)), ("child", "<u4", 4),
                  ("pad", "<u4", 4)])
TRI = np.dtype([("v0", "<f4", 3), ("id", "<u4"), ("v1", "<f4", 4), ("v2", "<f4", 4),
                ("pad", "<f4", 4)])


@pytest.fixture(scope="module", params=[(2, 1, 0), (4, 0, 0), (3, 1, 0), (2, 1, 2)],
                ids=["leaf2-area", "leaf4-grandchildren", "leaf3-area", "leaf2-area-treelets"])
def emu(request):
    """The emulated builder: leaves of <= 2 triangles + largest-area-first collapse (the device
    defaults), and the other settings of both knobs."""
    OUT.parent.mkdir(exist_ok=True)
    h = hashlib.sha256(SRC.read_bytes() + CORE.read_bytes()).hexdigest()
    stamp = OUT.with_suffix(".stamp")
    if not (OUT.exists() and stamp.exists() and stamp.read_text() == h):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", str(SRC), "-o",
                        str(OUT)], check=True)
        stamp.write_text(h)
    lib = C.CDLL(str(OUT))
    lib.lbvh_emu_morton.restype = C.c_uint64
    lib.lbvh_emu_morton.argtypes = [C.c_float] * 3
    lib.lbvh_emu_set_max_leaf(C.c_uint32(request.param[0]))
    lib.lbvh_emu_set_collapse_by_area(C.c_uint32(request.param[1]))
    lib.lbvh_emu_set_treelets(C.c_uint32(request.param[2]), C.c_uint32(7))
    lib.max_leaf = request.param[0]
    lib.treelets = request.param[2]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def build_blas(emu, scene, base2=3, base4=5):
    """Every BLAS of `scene` through the emulated device build."""
    entries = scene.array(_ffi.SCENE_ENTRIES)
    vertices = np.ascontiguousarray(scene.array(_ffi.SCENE_VERTICES))
    indices = np.ascontiguousarray(scene.array(_ffi.SCENE_INDICES))
    n_prims = int(entries["primitive_offset"][-1] + entries["primitive_count"][-1])
    counts = np.ascontiguousarray(entries["primitive_count"])
    prim_base = np.ascontiguousarray(entries["primitive_offset"])
    voff = np.ascontiguousarray(entries["vertex_offset"])
    ioff = np.ascontiguousarray(entries["index_offset"])
    n_seg = len(entries)
    cap = n_prims + 16
    nodes2, nodes4 = np.zeros(cap, NODE2), np.zeros(cap, NODE4)
    tris = np.zeros(n_prims, TRI)
    root2, root4 = np.zeros(n_seg, np.uint32), np.zeros(n_seg, np.uint32)
    root_box = np.zeros((n_seg, 6), np.float32)
    out = np.zeros(4, np.uint32)
    rc = emu.lbvh_emu_build_blas(_p(vertices), _p(indices), _p(counts), _p(prim_base), _p(voff),
                                 _p(ioff), C.c_uint32(n_seg), C.c_uint32(base2), C.c_uint32(base4),
                                 _p(nodes2), C.c_uint32(cap), _p(nodes4), C.c_uint32(cap), _p(tris),
                                 _p(root2), _p(root4), _p(root_box), _p(out))
    return dict(rc=rc, nodes2=nodes2, nodes4=nodes4, tris=tris, root2=root2, root4=root4,
                max_leaf=emu.max_leaf,
                root_box=root_box, n2=int(out[0]), n4=int(out[1]), depth4=int(out[2]),
                depth2=int(out[3]),
                entries=entries, vertices=vertices, indices=indices, base2=base2, base4=base4)


def _tri_bounds(tris, first, count):
    pts = np.concatenate([tris["v0"][first:first + count], tris["v1"][first:first + count, :3],
                          tris["v2"][first:first + count, :3]])
    return pts.min(0), pts.max(0)


def walk4(b, ref, seen, depth=1):
    """(lo, hi, depth) of the subtree behind child reference `ref` of the 4-wide array."""
    if ref & LEAF:
        count, first = ((ref >> 28) & 7) + 1, ref & 0x0FFFFFFF
        assert count <= b.get("max_leaf", 4)
        seen.extend(range(first, first + count))
        return (*_tri_bounds(b["tris"], first, count), depth - 1)
    assert b["base4"] <= ref < b["base4"] + b["n4"], ref
    node = b["nodes4"][ref]
    lo, hi, deepest, used = [], [], depth, 0
    for s in range(4):
        c = int(node["child"][s])
        if c == NONE:
            assert np.all(np.isposinf(node["lo"][:, s])) and np.all(np.isneginf(node["hi"][:, s]))
            continue
        used += 1
        l, h, d = walk4(b, c, seen, depth + 1)
        assert np.array_equal(node["lo"][:, s], l) and np.array_equal(node["hi"][:, s], h), ref
        lo.append(l), hi.append(h)
        deepest = max(deepest, d)
    assert used >= 2
    return np.min(lo, 0), np.max(hi, 0), deepest


def walk2(b, ref, seen, depth=1):
    """(lo, hi, interior levels) of the subtree behind `ref` of the 2-wide array."""
    if ref & LEAF:
        count, first = ((ref >> 28) & 7) + 1, ref & 0x0FFFFFFF
        seen.extend(range(first, first + count))
        return (*_tri_bounds(b["tris"], first, count), depth - 1)
    assert b["base2"] <= ref < b["base2"] + b["n2"], ref
    node = b["nodes2"][ref]
    q = node["q"]
    boxes = [(q[0:3], q[3:6]), (q[6:9], q[9:12])]
    lo, hi, deepest = [], [], depth
    for s in range(2):
        l, h, d = walk2(b, int(node["child"][s]), seen, depth + 1)
        assert np.array_equal(boxes[s][0], l) and np.array_equal(boxes[s][1], h)
        lo.append(l), hi.append(h)
        deepest = max(deepest, d)
    return np.min(lo, 0), np.max(hi, 0), deepest


def check_blas(b):
    """`b` = the builder's outputs (emulated, or read back from the device: then root_box /
    depth4 are None and only the BLASes some instance uses have a known root)."""
    assert b["rc"] == 0
    deepest = deepest2 = 0
    for e, ent in enumerate(b["entries"]):
        n, base = int(ent["primitive_count"]), int(ent["primitive_offset"])
        if b.get("known_roots") is not None and e not in b["known_roots"]:
            continue
        if n == 0:
            assert b["root2"][e] == NONE and b["root4"][e] == NONE
            continue
        for walk, root in ((walk4, b["root4"]), (walk2, b["root2"])):
            seen = []
            res = walk(b, int(root[e]), seen)
            assert sorted(seen) == list(range(base, base + n)), "every triangle exactly once"
            if b.get("root_box") is not None:
                assert np.array_equal(res[0], b["root_box"][e, :3])
                assert np.array_equal(res[1], b["root_box"][e, 3:])
            if walk is walk4:
                deepest = max(deepest, res[2])
            else:
                deepest2 = max(deepest2, res[2])
        # the triangle records are the BLAS's triangles, each once, with their original index
        t = b["tris"][base:base + n]
        assert sorted(t["id"].tolist()) == list(range(n))
        ix = b["indices"][int(ent["index_offset"]):int(ent["index_offset"]) + 3 * n].reshape(n, 3)
        pos = b["vertices"]["position"][int(ent["vertex_offset"]):]
        assert np.array_equal(t["v0"], pos[ix[t["id"], 0]])
        assert np.array_equal(t["v1"][:, :3], pos[ix[t["id"], 1]])
        assert np.array_equal(t["v2"][:, :3], pos[ix[t["id"], 2]])
        assert not t["v1"][:, 3].any() and not t["v2"][:, 3].any() and not t["pad"].any()
    if b.get("depth4") is not None:
        assert deepest == b["depth4"]
        # the reported 2-wide depth counts the radix tree's interior nodes above a primitive,
        # including those folded into a leaf (a chain of at most max_leaf - 1): never less than
        # the emitted tree's depth
        assert deepest2 <= b["depth2"] <= deepest2 + b["max_leaf"] - 1
    return deepest


def test_morton_code_orders_like_interleaved_bits(emu):
    rng = np.random.default_rng(3)
    for x, y, z in rng.random((200, 3)):
        code = emu.lbvh_emu_morton(x, y, z)
        q = [min(int(np.float32(v) * np.float32(2097152.0)), 2097151) for v in (x, y, z)]
        want = 0
        for bit in range(21):
            for a in range(3):
                want |= ((q[a] >> bit) & 1) << (3 * bit + 2 - a)
        assert code == want
    assert emu.lbvh_emu_morton(1.0, 1.0, 1.0) == (1 << 63) - 1
    assert emu.lbvh_emu_morton(-1.0, float("nan"), 0.0) == 0


def test_cornell_box_every_blas_is_a_single_leaf_or_tiny_tree(emu):
    c = scenes.cornell_box()
    b = build_blas(emu, c["scene"])
    check_blas(b)
    # 5 meshes of 2..12 triangles; <= max_leaf triangles => the root is a leaf
    for e, ent in enumerate(b["entries"]):
        if 0 < ent["primitive_count"] <= emu.max_leaf:
            assert b["root4"][e] & LEAF and b["root4"][e] == b["root2"][e]


def test_icosphere_tree_is_valid_and_shallow(emu):
    pos, idx = scenes.icosphere(4)  # 5120 triangles
    s = lb.Scene()
    s.blas.add_bvh_indexed(pos, idx)
    s.blas.add_bvh_indexed(pos * 0.25 + 3.0, idx[: 3 * 777])
    b = build_blas(emu, s)
    depth = check_blas(b)
    assert 4 <= depth <= 14, depth            # ~log4(5120 / leaf) + slack for the Morton splits
    assert b["n4"] < b["n2"] <= 5120 + 777    # the collapse removes levels
    # a tree with leaves of <= L triangles has >= n/L leaves => >= n/L - 1 interior nodes
    assert b["n2"] >= (5120 + 777) // emu.max_leaf - 2


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 9, 33])
def test_small_counts(emu, n):
    rng = np.random.default_rng(n)
    s = lb.Scene()
    s.blas.add_bvh(rng.random((3 * n, 3), np.float32) * 4 - 2)
    b = build_blas(emu, s, base2=0, base4=0)
    check_blas(b)
    if n <= emu.max_leaf:
        assert b["n2"] == 0 and b["n4"] == 0 and b["depth4"] == 0


def test_duplicate_and_degenerate_triangles(emu):
    """Equal Morton codes (identical triangles, zero-extent bounds) are told apart by their
    sorted position; the tree stays balanced instead of degenerating into a list."""
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    s = lb.Scene()
    s.blas.add_bvh(np.tile(tri, (100, 1)))                      # 100 copies of one triangle
    s.blas.add_bvh(np.zeros((3 * 37, 3), np.float32) + 2.5)     # 37 points: zero extent
    pts = np.random.default_rng(5).random((3 * 64, 3), np.float32)
    pts[:, 1] = 0.5                                             # flat: one axis has no extent
    s.blas.add_bvh(pts)
    b = build_blas(emu, s)
    depth = check_blas(b)
    assert depth <= 5  # 100 equal codes: positions split evenly => depth ~ log4(100 / leaf) + 1


def test_clustered_soup_stays_within_the_stack_limit(emu):
    """Widely separated tight clusters (deep Morton prefixes) must still fit the traversal
    stack (31 4-wide levels)."""
    rng = np.random.default_rng(11)
    centres = np.array([[0, 0, 0], [1e3, 0, 0], [1e3, 1e-3, 0], [-5e2, 7e2, 3e2]], np.float32)
    tris = []
    for c in centres:
        p = c + rng.normal(0, 1e-4, (3 * 400, 3)).astype(np.float32)
        tris.append(p)
    s = lb.Scene()
    s.blas.add_bvh(np.concatenate(tris))
    b = build_blas(emu, s)
    depth = check_blas(b)
    assert depth <= 31


def _gpu_instances(scene):
    """128-byte instance records as the library lays them out (host copy)."""
    scene_gpu = scene.array(_ffi.SCENE_GPU_INSTANCES)
    return np.ascontiguousarray(scene_gpu)


def test_tlas_over_instances(emu):
    c = scenes.spheres_1m(grid=3, subdivisions=1)  # 9 spheres + ground
    scene = c["scene"]
    b = build_blas(emu, scene)
    inst = _gpu_instances(scene)
    entries = b["entries"]
    ids = np.array([i for i in range(len(inst)) if entries["primitive_count"][inst["blas"][i]] > 0],
                   np.uint32)
    n = len(ids)
    nodes2, nodes4 = np.zeros(n + 1, NODE2), np.zeros(n + 1, NODE4)
    root2, root4, out = np.zeros(1, np.uint32), np.zeros(1, np.uint32), np.zeros(4, np.uint32)
    blas_of = np.ascontiguousarray(inst["blas"])
    rc = emu.lbvh_emu_build_tlas(_p(inst), _p(blas_of), _p(b["root_box"]), _p(ids), C.c_uint32(n),
                                 _p(nodes2), C.c_uint32(n + 1), _p(nodes4), C.c_uint32(n + 1),
                                 _p(root2), _p(root4), _p(out))
    assert rc == 0
    n2, n4 = int(out[0]), int(out[1])
    assert n2 == n - 1 and 1 <= n4 <= n2  # leaves of one instance: a full binary tree

    def world_box(i):
        rb = b["root_box"][inst["blas"][i]]
        m = inst["o2w"][i].reshape(3, 4)
        corners = np.array([[rb[3 * (k & 1)], rb[1 + 3 * ((k >> 1) & 1)], rb[2 + 3 * ((k >> 2) & 1)], 1]
                            for k in range(8)], np.float32)
        w = corners @ m.T
        return w.min(0), w.max(0)

    def walk(ref, seen):
        if ref & LEAF:
            i = ref & 0x0FFFFFFF
            seen.append(i)
            return world_box(i)
        node = nodes4[ref]
        lo, hi = [], []
        for s in range(4):
            ch = int(node["child"][s])
            if ch == NONE:
                continue
            l, h = walk(ch, seen)
            if ch & LEAF:   # padded by 4 ulp like Scene::build_tlas; never smaller than the box
                assert np.all(node["lo"][:, s] <= l) and np.all(node["hi"][:, s] >= h)
                assert np.allclose(node["lo"][:, s], l, rtol=1e-5, atol=1e-5)
                l, h = node["lo"][:, s], node["hi"][:, s]
            else:
                assert np.array_equal(node["lo"][:, s], l) and np.array_equal(node["hi"][:, s], h)
            lo.append(l), hi.append(h)
        return np.min(lo, 0), np.max(hi, 0)

    seen = []
    walk(int(root4[0]), seen)
    assert sorted(seen) == sorted(ids.tolist())
    # the host TLAS covers the same instances
    host = scene.array(_ffi.SCENE_TLAS_NODES)
    assert sorted(host["left_first"][host["count"] > 0].tolist()) == sorted(ids.tolist())


def test_empty_scene_and_single_instance(emu):
    s = lb.Scene()
    b = build_blas(emu, s)
    assert b["rc"] == 0 and b["n2"] == 0 and b["n4"] == 0
    assert b["root4"][0] == NONE
    root2, root4, out = np.zeros(1, np.uint32), np.zeros(1, np.uint32), np.zeros(4, np.uint32)
    z = np.zeros(64, np.float32)
    rc = emu.lbvh_emu_build_tlas(_p(z), _p(z), _p(z), _p(z), C.c_uint32(0), _p(z), C.c_uint32(1),
                                 _p(z), C.c_uint32(1), _p(root2), _p(root4), _p(out))
    assert rc == 0 and root4[0] == NONE and root2[0] == NONE


def _sah2(b, ref, ci=1.2, ct=1.0):
    """Surface-area cost of the 2-wide tree behind `ref` (C_i per interior node, C_t per
    triangle, both weighted by the half area of the node's box)."""
    def area(lo, hi):
        d = np.maximum(hi.astype(np.float64) - lo, 0)
        return d[0] * d[1] + d[1] * d[2] + d[2] * d[0]

    def rec(ref, lo, hi):
        if ref & LEAF:
            return ct * area(lo, hi) * (((ref >> 28) & 7) + 1)
        q = b["nodes2"][ref]["q"]
        ch = b["nodes2"][ref]["child"]
        return ci * area(lo, hi) + rec(int(ch[0]), q[0:3], q[3:6]) + rec(int(ch[1]), q[6:9], q[9:12])

    q = b["nodes2"][ref]["q"]
    lo, hi = np.minimum(q[0:3], q[6:9]), np.maximum(q[3:6], q[9:12])
    return rec(ref, lo, hi)


@pytest.mark.parametrize("n_tris,seed", [(5, 0), (6, 1), (7, 2), (7, 3)])
def test_treelet_restructuring_finds_the_optimal_small_tree(emu, n_tris, seed):
    """A BLAS of <= 7 triangles with single-triangle leaves IS one treelet: after one pass its
    surface-area cost must equal the minimum over ALL binary trees on those leaves, computed
    here by an independent exhaustive search (checks the subset DP and its partition walk)."""
    from itertools import combinations
    rng = np.random.default_rng(seed)
    pos = (rng.random((n_tris, 1, 3)) * 4 + rng.normal(0, 0.3, (n_tris, 3, 3))).astype(np.float32)
    s = lb.Scene()
    s.blas.add_bvh(pos.reshape(-1, 3))
    emu.lbvh_emu_set_max_leaf(C.c_uint32(1))
    emu.lbvh_emu_set_treelets(C.c_uint32(1), C.c_uint32(2))
    try:
        b = build_blas(emu, s, base2=0, base4=0)
        b["max_leaf"] = 1
        check_blas(b)
        emu.lbvh_emu_set_treelets(C.c_uint32(0), C.c_uint32(7))
        plain = build_blas(emu, s, base2=0, base4=0)
    finally:
        emu.lbvh_emu_set_max_leaf(C.c_uint32(emu.max_leaf))
        emu.lbvh_emu_set_treelets(C.c_uint32(emu.treelets), C.c_uint32(7))
    got, before = _sah2(b, int(b["root2"][1])), _sah2(plain, int(plain["root2"][1]))

    lo, hi = pos.min(1).astype(np.float64), pos.max(1).astype(np.float64)

    def area(idx):
        d = hi[list(idx)].max(0) - lo[list(idx)].min(0)
        return d[0] * d[1] + d[1] * d[2] + d[2] * d[0]

    best = {}

    def opt(sub):
        if len(sub) == 1:
            return area(sub)
        if sub not in best:
            items = sorted(sub)
            rest = items[1:]
            cands = []
            for k in range(len(rest)):
                for pick in combinations(rest, k):
                    left = frozenset([items[0], *pick])
                    cands.append(opt(left) + opt(sub - left))
            best[sub] = 1.2 * area(sub) + min(cands)
        return best[sub]

    want = opt(frozenset(range(n_tris)))
    assert got == pytest.approx(want, rel=1e-5)
    assert got <= before * (1 + 1e-6)


def test_treelet_passes_lower_the_cost_of_a_real_mesh(emu):
    """Two passes over a displaced icosphere: the 2-wide surface-area cost falls by > 3 % and
    the tree stays valid (check_blas) -- the quantity the trace rate follows."""
    c = scenes.spheres_1m(grid=1, subdivisions=4)
    costs = []
    try:
        for passes in (0, 2):
            emu.lbvh_emu_set_treelets(C.c_uint32(passes), C.c_uint32(7))
            b = build_blas(emu, c["scene"], base2=0, base4=0)
            check_blas(b)
            costs.append(_sah2(b, int(b["root2"][1])))
    finally:
        emu.lbvh_emu_set_treelets(C.c_uint32(emu.treelets), C.c_uint32(7))
    assert costs[1] < 0.97 * costs[0], costs
