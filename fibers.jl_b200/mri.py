"""Minimal `MRI` container mirroring the fields of the reference struct that the
reconstruction path reads or writes (reference: src/mri.jl:80-130 struct, :249-265 ctor).

Volume I/O (mri_read / mri_write, NIfTI / MGH) stays on the host and is out of scope
(SURVEY.md §2 row 7); this class only carries the layout contract: `vol` is a Julia
column-major `[nx, ny, nz, nframes]` array (numpy order='F'), `bval` float32 [nframes],
`bvec` float32 [nframes, 3].
"""
from __future__ import annotations

import numpy as np


class MRI:
    def __init__(self, vol=None, bval=None, bvec=None, **header):
        self.vol = None if vol is None else np.asfortranarray(vol)
        self.bval = np.zeros(0, np.float32) if bval is None else np.ascontiguousarray(bval, np.float32)
        self.bvec = np.zeros((0, 3), np.float32) if bvec is None else np.asfortranarray(bvec, np.float32)
        self.header = dict(header)

    @property
    def volsize(self):
        return tuple(self.vol.shape[:3])

    @property
    def nframes(self):
        return 1 if self.vol.ndim < 4 else self.vol.shape[3]

    @classmethod
    def like(cls, ref: "MRI", nframes: int = 1, dtype=np.float32) -> "MRI":
        """MRI(ref, nframes, datatype): zero-filled, header copied from `ref`, 3-D when
        nframes == 1 (reference: src/mri.jl:249-265)."""
        shp = ref.volsize + ((nframes,) if nframes > 1 else ())
        return cls(np.zeros(shp, dtype, order="F"), **ref.header)
