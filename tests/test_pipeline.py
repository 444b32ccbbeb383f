"""The reference's tutorial flow end to end on the GPU path: mri_read (NIfTI + b-tables, pinned) -> dti_fit / gqi_rec ->
mri_write of the maps -> mri_read back -> stream on the peaks (docs/tutorial.ipynb: mri_read :3940, gqi_rec :4674, stream :7265)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


@pytest.mark.gpu
def test_read_recon_write_stream(tmp_path):
    import fibers_jl_b200 as Fb
    import fibers_oracle as O
    import stream_oracle as SO
    from fibers_jl_b200 import phantom
    ph = phantom.gqi_phantom((14, 12, 8), seed=9, mask_fill=0.8)
    M = np.array([[-2, 0, 0, 14], [0, 2, 0, -12], [0, 0, 2, -8], [0, 0, 0, 1]], np.float32)
    dwi = Fb.MRI(ph["dwi"], ph["bval"], ph["bvec"], vox2ras0=M, tr=3000.0)
    Fb.mri_write(dwi, str(tmp_path / "dwi.nii.gz"))
    Fb.mri_write(Fb.MRI(ph["mask"].astype(np.uint8), vox2ras0=M), str(tmp_path / "mask.mgz"))
    d = Fb.mri_read(str(tmp_path / "dwi.nii.gz"), pin=True)        # page-locked: the recon call DMAs straight from it
    m = Fb.mri_read(str(tmp_path / "mask"))                        # stem: finds mask.mgz
    try:
        assert d.vol.dtype == np.float32 and np.array_equal(d.vol, ph["dwi"]) and np.array_equal(m.vol, ph["mask"].astype(np.uint8))
        np.testing.assert_array_equal(d.bval, ph["bval"])
        got = Fb.gqi_rec(d, m)
        ref = Fb.gqi_rec(Fb.MRI(ph["dwi"], ph["bval"], d.bvec), Fb.MRI(ph["mask"]))
        assert np.array_equal(got.odf.vol, ref.odf.vol) and all(np.array_equal(a.vol, b.vol) for a, b in zip(got.peak, ref.peak))
    finally:
        Fb._lib.lib().fibers_cuda_host_unregister(Fb._lib.ptr(d.vol))
    for k, pk in enumerate(got.peak):
        pk.header.update(vox2ras0=M)
        Fb.mri_write(pk, str(tmp_path / f"peak{k + 1}.nii.gz"))
    peaks = [Fb.mri_read(str(tmp_path / f"peak{k + 1}.nii.gz")) for k in range(3)]
    assert all(np.array_equal(a.vol, b.vol) for a, b in zip(peaks, got.peak))
    sub = Fb.draw_sublist(2, rng=3)
    tr = Fb.stream(peaks, f=got.qa, f_thresh=0.02, mask=m, sublist=sub)
    want = SO.stream([p.vol for p in got.peak], list(sub), f=[q.vol for q in got.qa], f_thresh=0.02, mask=ph["mask"])
    assert tr.n_count == len(want) > 0 and all(np.array_equal(a, b) for a, b in zip(tr.xyz, want))
    # ... and to disk (trk_write, src/trk.jl:433): header from the mask volume, points as (xyz + .5) * voxel_size
    assert Fb.trk_write(tr, str(tmp_path / "tract.trk")) is False
    raw = (tmp_path / "tract.trk").read_bytes()
    import struct
    assert len(raw) == 1000 + 4 * tr.n_count + 12 * int(tr.npts.sum()) and struct.unpack_from("<3h", raw, 6) == (14, 12, 8)
    assert struct.unpack_from("<i", raw, 988)[0] == tr.n_count and raw[948:951] == b"LAS"
