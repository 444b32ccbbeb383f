#!/usr/bin/env python
"""Export the ODF tessellation tables of the reference to a binary fixture.

Reads the literal tables `sphere_362`, `sphere_642`, `sphere_724` from the
reference's src/odf.jl (reference: src/odf.jl:14-1100, :1104-3030, :3034-5206)
and writes `fibers.jl_b200/data/spheres.npz` with, per sphere,
  vertices_<n>  float32 [2M, 3]   (Float64 literal -> Float32, as `Float32.([...])` does)
  faces_<n>     int32   [F, 3]    (1-based, exactly as in the reference)

These are DATA (direction tables of DTK / DSI Studio), not code.  Peak outputs of
gqi_rec / dsi_rec are verbatim copies of vertex rows (src/gqi.jl:154-155), so the
table must be bit-identical.  Run only in the build container (needs /root/reference).
"""
import re
import sys
import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/odf.jl"
OUT = sys.argv[2] if len(sys.argv) > 2 else "fibers.jl_b200/data/spheres.npz"


def main():
    lines = open(REF).read().split("\n")
    starts = [(i, re.search(r"sphere_(\d+)\s*=\s*ODF\(", l).group(1))
              for i, l in enumerate(lines) if re.search(r"const global sphere_\d+\s*=\s*ODF\(", l)]
    out = {}
    for (i0, name) in starts:
        # vertices block: from "Float32.([" to "]),"; faces block: from "[" to "])"
        i = i0
        while "Float32.([" not in lines[i]:
            i += 1
        i += 1
        verts = []
        while lines[i].strip() != "]),":
            t = lines[i].split()
            if t:
                assert len(t) == 3, (i, lines[i])
                verts.append([np.float32(float(x)) for x in t])
            i += 1
        i += 1
        assert lines[i].strip() == "[", (i, lines[i])
        i += 1
        faces = []
        while lines[i].strip() != "])":
            t = lines[i].split()
            if t:
                assert len(t) == 3
                faces.append([int(x) for x in t])
            i += 1
        v = np.asarray(verts, dtype=np.float32)
        f = np.asarray(faces, dtype=np.int32)
        assert v.shape[0] == int(name), (name, v.shape)
        assert f.min() == 1 and f.max() == v.shape[0]
        out["vertices_" + name] = v
        out["faces_" + name] = f
        print(name, v.shape, f.shape)
    np.savez_compressed(OUT, **out)


if __name__ == "__main__":
    main()
