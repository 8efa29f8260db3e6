"""Minimal GLB writer for the loader tests: one textured quad per mesh, images embedded as
bufferViews (PNG / JPEG file bytes from tests/golden/image_fixtures.npz)."""
import json
import struct

import numpy as np


def textured_quad_glb(image_files, texture_sources, materials, mime="image/png"):
    """image_files: list of encoded image bytes; texture_sources: textures[i].source;
    materials: list of dicts {base: texture index or None, mr: texture index or None,
    factor: rgba, rough: float, metal: float}.  One quad mesh + node per material, side by
    side along +x, facing +z, uv = [0,1]^2 with v down."""
    blob = bytearray()
    views, accessors = [], []

    def add_view(data: bytes, target=None):
        while len(blob) % 4:
            blob.append(0)
        v = {"buffer": 0, "byteOffset": len(blob), "byteLength": len(data)}
        if target:
            v["target"] = target
        blob.extend(data)
        views.append(v)
        return len(views) - 1

    pos = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], dtype=np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], dtype=np.float32), (4, 1))
    uv = np.array([[0, 1], [1, 1], [1, 0], [0, 0]], dtype=np.float32)
    idx = np.array([0, 1, 2, 0, 2, 3], dtype=np.uint16)

    def add_accessor(arr, ctype, atype, target, minmax=False):
        a = {"bufferView": add_view(arr.tobytes(), target), "componentType": ctype,
             "count": int(arr.shape[0]), "type": atype}
        if minmax:
            a["min"], a["max"] = arr.min(0).tolist(), arr.max(0).tolist()
        accessors.append(a)
        return len(accessors) - 1

    a_pos = add_accessor(pos, 5126, "VEC3", 34962, True)
    a_nrm = add_accessor(nrm, 5126, "VEC3", 34962)
    a_uv = add_accessor(uv, 5126, "VEC2", 34962)
    a_idx = add_accessor(idx, 5123, "SCALAR", 34963)
    images = [{"bufferView": add_view(bytes(f)), "mimeType": mime} for f in image_files]
    textures = [{"source": s} for s in texture_sources]
    mats, meshes, nodes = [], [], []
    for k, m in enumerate(materials):
        pbr = {"baseColorFactor": list(m.get("factor", (1, 1, 1, 1))),
               "roughnessFactor": m.get("rough", 1.0), "metallicFactor": m.get("metal", 0.0)}
        if m.get("base") is not None:
            pbr["baseColorTexture"] = {"index": m["base"]}
        if m.get("mr") is not None:
            pbr["metallicRoughnessTexture"] = {"index": m["mr"]}
        mats.append({"pbrMetallicRoughness": pbr})
        meshes.append({"primitives": [{"attributes": {"POSITION": a_pos, "NORMAL": a_nrm,
                                                      "TEXCOORD_0": a_uv},
                                       "indices": a_idx, "material": k}]})
        nodes.append({"mesh": k, "translation": [2.2 * k, 0.0, 0.0]})
    while len(blob) % 4:
        blob.append(0)
    doc = {"asset": {"version": "2.0"}, "buffers": [{"byteLength": len(blob)}],
           "bufferViews": views, "accessors": accessors, "images": images, "textures": textures,
           "materials": mats, "meshes": meshes, "nodes": nodes,
           "scenes": [{"nodes": list(range(len(nodes)))}], "scene": 0}
    js = json.dumps(doc).encode()
    js += b" " * (-len(js) % 4)
    total = 12 + 8 + len(js) + 8 + len(blob)
    return (b"glTF" + struct.pack("<II", 2, total) + struct.pack("<I", len(js)) + b"JSON" + js +
            struct.pack("<I", len(blob)) + b"BIN\x00" + bytes(blob))
