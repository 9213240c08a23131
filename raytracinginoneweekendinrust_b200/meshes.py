"""OBJ ingestion with the reference's ``load_to_tris`` rules (src/main.rs:745-789) and seeded
stand-in meshes for the three OBJ models the reference ships only as git-LFS pointer stubs.

``load_to_tris``: tobj with ``triangulate: true``, **first model only**, positions only, one
flat ``Tri`` per face (OBJ normals and materials are ignored).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np


def load_obj_first_model(path: str) -> np.ndarray:
    """Returns (n, 9) float32 triangles of the first object/group, polygons fan-triangulated."""
    verts = []
    tris = []
    models_seen = 0
    faces_in_model = 0
    with open(path, "r", errors="replace") as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                verts.append([float(t[1]), float(t[2]), float(t[3])])
            elif t[0] in ("o", "g"):
                # tobj starts a new model at each o/g statement that follows faces
                if faces_in_model > 0:
                    models_seen += 1
                    if models_seen >= 1:
                        break
            elif t[0] == "f":
                idx = []
                for tok in t[1:]:
                    i = int(tok.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                for k in range(1, len(idx) - 1):
                    tris.append([idx[0], idx[k], idx[k + 1]])
                    faces_in_model += 1
    v = np.asarray(verts, np.float32)
    t = np.asarray(tris, np.int64)
    if len(t) == 0:
        return np.zeros((0, 9), np.float32)
    return v[t].reshape(-1, 9).astype(np.float32)


# approximate face counts and extents: bunny_2000_scale ~ a 5k-face bunny scaled to ~300 units;
# gargoyle / igea are 10^5-class scans (16.8 MB / 23.0 MB OBJ per the LFS pointers)
_STANDINS = {"bunny": (5000, 150.0, 11), "gargoyle": (200000, 160.0, 23), "igea": (268000, 160.0, 37)}


def synthesize(kind: str, n_tris: int | None = None) -> np.ndarray:
    """Seeded closed star-shaped surface (lat-long grid, radius modulated by a few low-frequency
    lobes), resting on y = 0 like the reference's models do in the Cornell box."""
    target, radius, seed = _STANDINS[kind]
    n = int(n_tris or target)
    rows = max(4, int(np.sqrt(n / 4.0)))
    cols = max(6, int(np.ceil(n / (2.0 * rows))))
    rs = np.random.RandomState(seed)
    theta = np.linspace(0.0, np.pi, rows + 1)           # polar
    phi = np.linspace(0.0, 2.0 * np.pi, cols + 1)[:-1]  # azimuth (wraps)
    T, P = np.meshgrid(theta, phi, indexing="ij")
    r = np.ones_like(T)
    for k in range(1, 6):
        a, b = rs.uniform(-1, 1, 2)
        r += 0.18 / k * np.sin(k * T * 2 + 3 * a) * np.cos(k * P + 3 * b)
    r *= radius
    x = r * np.sin(T) * np.cos(P)
    y = r * np.cos(T)
    z = r * np.sin(T) * np.sin(P)
    y = y - y.min()
    pts = np.stack([x, y, z], axis=-1).astype(np.float32)
    tris = []
    for i in range(rows):
        for j in range(cols):
            j2 = (j + 1) % cols
            a, b, c, d = pts[i, j], pts[i + 1, j], pts[i + 1, j2], pts[i, j2]
            if i != 0:
                tris.append(np.concatenate([a, b, d]))
            if i != rows - 1:
                tris.append(np.concatenate([b, c, d]))
    return np.asarray(tris, np.float32)


def write_obj(path: str, tris: np.ndarray) -> None:
    tris = np.asarray(tris, np.float32).reshape(-1, 3, 3)
    with open(path, "w") as f:
        f.write("o standin\n")
        for t in tris:
            for v in t:
                f.write(f"v {v[0]:.9g} {v[1]:.9g} {v[2]:.9g}\n")
        for i in range(len(tris)):
            f.write(f"f {3 * i + 1} {3 * i + 2} {3 * i + 3}\n")


def load_or_synthesize(obj_path: str | None, kind: str, n_tris: int | None = None) -> np.ndarray:
    if obj_path and Path(obj_path).exists() and Path(obj_path).stat().st_size > 1024:
        return load_obj_first_model(obj_path)
    return synthesize(kind, n_tris)
