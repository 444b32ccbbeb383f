"""`ODF` tessellation container and the three sphere tables of the reference
(reference: src/odf.jl:8-11 struct; :14 sphere_362, :1104 sphere_642 (default), :3034 sphere_724).

The tables are data exported bit-exactly from the reference by tools/export_spheres.py into
data/spheres.npz (peak outputs are verbatim copies of vertex rows, src/gqi.jl:154-155).
"""
from __future__ import annotations

import os

import numpy as np


class ODF:
    def __init__(self, vertices, faces):
        self.vertices = np.asfortranarray(vertices, np.float32)   # [2M, 3]
        self.faces = np.asfortranarray(faces, np.int32)           # [F, 3], 1-based

    @property
    def nvert(self) -> int:
        return self.vertices.shape[0] // 2


def _load(n: int) -> ODF:
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "spheres.npz"))
    return ODF(z[f"vertices_{n}"], z[f"faces_{n}"])


sphere_362 = _load(362)
sphere_642 = _load(642)
sphere_724 = _load(724)
