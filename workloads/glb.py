"""Minimal GLB (binary glTF 2.0) triangle extractor for the reference's test scenes
(data/test_assets/scenes/*.glb): all TRIANGLES primitives of the default scene, node transforms
applied.  Only what those files use: float32 POSITION, u8/u16/u32 indices, TRS or matrix nodes."""
from __future__ import annotations

import json
import struct

import numpy as np

_COMP = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


def _quat_to_mat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float64)


def _node_matrix(node):
    if "matrix" in node:
        return np.asarray(node["matrix"], np.float64).reshape(4, 4).T  # column-major in the file
    m = np.eye(4)
    r = _quat_to_mat(node.get("rotation", [0, 0, 0, 1]))
    s = np.asarray(node.get("scale", [1, 1, 1]), np.float64)
    m[:3, :3] = r * s[None, :]
    m[:3, 3] = node.get("translation", [0, 0, 0])
    return m


def load_triangles(path: str):
    """(verts [V,3] float32, tris [T,3] int32) of a .glb file."""
    data = open(path, "rb").read()
    magic, version, length = struct.unpack_from("<III", data, 0)
    assert magic == 0x46546C67 and version == 2, "not a GLB 2.0 file"
    off = 12
    gltf, blob = None, b""
    while off < length:
        clen, ctype = struct.unpack_from("<II", data, off)
        chunk = data[off + 8: off + 8 + clen]
        if ctype == 0x4E4F534A:
            gltf = json.loads(chunk.decode("utf8"))
        elif ctype == 0x004E4942:
            blob = chunk
        off += 8 + clen

    def accessor(i):
        a = gltf["accessors"][i]
        bv = gltf["bufferViews"][a["bufferView"]]
        dt = np.dtype(_COMP[a["componentType"]])
        nc = _NCOMP[a["type"]]
        start = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        stride = bv.get("byteStride", 0) or dt.itemsize * nc
        raw = np.frombuffer(blob, np.uint8, count=(a["count"] - 1) * stride + dt.itemsize * nc, offset=start)
        rows = np.lib.stride_tricks.as_strided(raw, shape=(a["count"], dt.itemsize * nc), strides=(stride, 1))
        return np.ascontiguousarray(rows).view(dt).reshape(a["count"], nc)

    verts, tris = [], []
    base = 0

    def visit(ni, parent):
        nonlocal base
        node = gltf["nodes"][ni]
        m = parent @ _node_matrix(node)
        if "mesh" in node:
            for prim in gltf["meshes"][node["mesh"]]["primitives"]:
                if prim.get("mode", 4) != 4:
                    continue
                p = accessor(prim["attributes"]["POSITION"]).astype(np.float64)
                p = p @ m[:3, :3].T + m[:3, 3]
                idx = (accessor(prim["indices"]).reshape(-1) if "indices" in prim
                       else np.arange(len(p))).astype(np.int64)
                verts.append(p.astype(np.float32))
                tris.append((idx.reshape(-1, 3) + base).astype(np.int32))
                base += len(p)
        for c in node.get("children", []):
            visit(c, m)

    scene = gltf["scenes"][gltf.get("scene", 0)]
    for ni in scene["nodes"]:
        visit(ni, np.eye(4))
    return np.concatenate(verts), np.concatenate(tris)
