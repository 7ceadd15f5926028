"""Minimal binary glTF (GLB) writer for vertex-coloured triangle meshes.

The reference writes ``wing_pred.glb`` / ``wing_error.glb`` with trimesh + matplotlib
(``scripts/inference_gnn.py:62-221``); neither is needed for a positions + indices + COLOR_0 mesh, which is
all the Gradio viewer (``app.py``) loads.  Colour maps: the reference's 9-stop viridis table
(``inference_gnn.py:62-89``) and matplotlib's "hot" ramp (``:204-206``).
"""
from __future__ import annotations

import json
import struct
from pathlib import Path

import numpy as np

_VIRIDIS = np.array([[0.0, 68, 1, 84], [0.125, 72, 40, 120], [0.25, 62, 74, 137], [0.375, 49, 104, 142], [0.5, 38, 130, 142],
                     [0.625, 31, 158, 137], [0.75, 53, 183, 121], [0.875, 109, 205, 89], [1.0, 253, 231, 37]], dtype=np.float64)


def viridis_rgb(x: np.ndarray) -> np.ndarray:
    t = np.clip(np.asarray(x, dtype=np.float64).ravel(), 0.0, 1.0)
    out = np.empty((t.size, 3), dtype=np.uint8)
    for c in range(3):
        out[:, c] = np.clip(np.interp(t, _VIRIDIS[:, 0], _VIRIDIS[:, c + 1]), 0, 255).astype(np.uint8)
    return out


def hot_rgb(x: np.ndarray) -> np.ndarray:
    """matplotlib 'hot': red ramps over [0, 0.365], green over [0.365, 0.746], blue over [0.746, 1]."""
    t = np.clip(np.asarray(x, dtype=np.float64).ravel(), 0.0, 1.0)
    r = np.interp(t, [0.0, 0.365079, 1.0], [0.0416, 1.0, 1.0])
    g = np.interp(t, [0.0, 0.365079, 0.746032, 1.0], [0.0, 0.0, 1.0, 1.0])
    b = np.interp(t, [0.0, 0.746032, 1.0], [0.0, 0.0, 1.0])
    return (np.stack([r, g, b], axis=1) * 255).astype(np.uint8)


def write_glb(path: Path, vertices: np.ndarray, faces: np.ndarray, rgb: np.ndarray) -> None:
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    f = np.ascontiguousarray(faces, dtype=np.uint32).ravel()
    c = np.ascontiguousarray(np.hstack([rgb, np.full((len(rgb), 1), 255, np.uint8)]), dtype=np.uint8)

    def pad4(b: bytes) -> bytes:
        return b + b"\x00" * (-len(b) % 4)

    chunks = [pad4(f.tobytes()), pad4(v.tobytes()), pad4(c.tobytes())]
    offs = np.cumsum([0] + [len(b) for b in chunks])
    gltf = {
        "asset": {"version": "2.0", "generator": "deep_fem_uav_wing.gnn.glb"},
        "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 1, "COLOR_0": 2}, "indices": 0, "mode": 4}]}],
        "buffers": [{"byteLength": int(offs[-1])}],
        "bufferViews": [
            {"buffer": 0, "byteOffset": int(offs[0]), "byteLength": int(f.nbytes), "target": 34963},
            {"buffer": 0, "byteOffset": int(offs[1]), "byteLength": int(v.nbytes), "target": 34962},
            {"buffer": 0, "byteOffset": int(offs[2]), "byteLength": int(c.nbytes), "target": 34962},
        ],
        "accessors": [
            {"bufferView": 0, "componentType": 5125, "count": int(f.size), "type": "SCALAR"},
            {"bufferView": 1, "componentType": 5126, "count": int(len(v)), "type": "VEC3", "min": v.min(0).tolist() if len(v) else [0, 0, 0],
             "max": v.max(0).tolist() if len(v) else [0, 0, 0]},
            {"bufferView": 2, "componentType": 5121, "count": int(len(c)), "type": "VEC4", "normalized": True},
        ],
    }
    js = json.dumps(gltf, separators=(",", ":")).encode()
    js += b" " * (-len(js) % 4)
    bin_ = b"".join(chunks)
    total = 12 + 8 + len(js) + 8 + len(bin_)
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "wb") as fh:
        fh.write(struct.pack("<4sII", b"glTF", 2, total))
        fh.write(struct.pack("<I4s", len(js), b"JSON") + js)
        fh.write(struct.pack("<I4s", len(bin_), b"BIN\x00") + bin_)
