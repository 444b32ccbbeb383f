"""The pinning kit (tools/pin/*, julia/pin_golden.jl): anyone with Julia runs the real reference on the golden fixtures' inputs
and compares its volumes with the oracle's.  Julia is not available here, so this test exercises the two Python ends: the exported
inputs carry the fixtures bit for bit, and the checker accepts outputs equal to the oracle's and rejects perturbed ones."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
import fibers_jl_b200 as Fb  # noqa: E402
import fibers_oracle as O  # noqa: E402


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", "pin", name + ".py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def test_pin_kit_round_trip(tmp_path, capsys):
    exp, chk = _load("export_inputs"), _load("check_outputs")
    ind, outd = str(tmp_path / "in"), str(tmp_path / "out")
    exp.main(ind)
    os.makedirs(outd)
    verts = np.asarray(O.load_sphere(642)[0][:321], np.float32)
    for name in exp.FIXTURES:
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        # inputs: bit for bit, tables as float32 text that parses back exactly
        assert np.array_equal(Fb.mri_read(os.path.join(ind, name + "_dwi.nii.gz")).vol, g["dwi"])
        assert np.array_equal(np.loadtxt(os.path.join(ind, name + "_bval.txt"), dtype=np.float32), g["bval"])
        assert np.array_equal(np.loadtxt(os.path.join(ind, name + "_bvec.txt"), dtype=np.float32).reshape(-1, 3), g["bvec"])
        assert np.array_equal(Fb.mri_read(os.path.join(ind, name + "_mask.nii.gz")).vol.reshape(g["mask"].shape) > 0, g["mask"] > 0)

        def put(field, arr):
            Fb.mri_write(Fb.MRI(np.asfortranarray(np.asarray(arr, np.float32))), os.path.join(outd, f"{name}_{field}.nii.gz"))
        if name == "rumba_small":                            # what rumba_write would leave (SNR estimates as text)
            for f in ("fodf", "fgm", "fcsf", "gfa", "var"):
                put(f, g[f])
            for k in range(5):
                put(f"peak{k + 1}", g["peak"][k])
            for f in ("snr_mean", "snr_std"):
                with open(os.path.join(outd, f"{name}_{f}.txt"), "w") as fh:
                    fh.write(f"{np.float32(g[f])}\n")
        elif name == "dti_small":                            # what dti_write / mri_write of the reference would leave
            for f in ("s0", "eigval1", "eigval2", "eigval3", "eigvec1", "rd", "md", "fa", "adc", "adc_s0"):
                put(f, g[f])
        else:
            put("odf", g["odf"])
            if "pdf" in g.files:
                put("pdf", g["pdf"])
            for k in range(3):
                idx = g["peak_idx"][..., k]
                put(f"peak{k + 1}", np.where((idx >= 0)[..., None], verts[np.maximum(idx, 0)], 0))
                put(f"qa{k + 1}", g["qa"][k])
    # the Tract of stream() as the reference's trk_write would leave it
    from fibers_jl_b200.stream import Tract, trk_write
    g = np.load(os.path.join(ROOT, "tests", "golden", "stream_small.npz"))
    assert np.array_equal(Fb.mri_read(os.path.join(ind, "stream_small_ovec2.nii.gz")).vol, g["ovec"][1])
    ends = np.cumsum(g["npts"])
    lines = [np.asfortranarray(g["xyz"][:, e - n:e]) for e, n in zip(ends, g["npts"])]
    M = np.diag([2.0, 2.0, 2.0, 1.0]).astype(np.float32)
    trk_write(Tract(lines, g["npts"], None, dict(volsize=list(g["mask"].shape), volres=[2.0, 2.0, 2.0], vox2ras0=M)), os.path.join(outd, "stream_small.trk"))
    # st_recon: eigenvalues as 3 frames, eigenvectors as 9 frames
    g = np.load(os.path.join(ROOT, "tests", "golden", "structens_small.npz"))
    assert np.array_equal(Fb.mri_read(os.path.join(ind, "structens_small_vol.nii.gz")).vol.reshape(g["vol"].shape), g["vol"])
    Fb.mri_write(Fb.MRI(np.asfortranarray(g["eigval"].astype(np.float32))), os.path.join(outd, "structens_small_eigval.nii.gz"))
    Fb.mri_write(Fb.MRI(np.asfortranarray(g["eigvec"].reshape(g["vol"].shape + (9,), order="F"))), os.path.join(outd, "structens_small_eigvec.nii.gz"))
    assert chk.check(outd) == 0
    assert "PINNED" in capsys.readouterr().out
    # a reference that disagreed would be caught: 0.1 % on the GQI ODF, a swapped DSI peak
    g = np.load(os.path.join(ROOT, "tests", "golden", "gqi_small.npz"))
    Fb.mri_write(Fb.MRI(np.asfortranarray((g["odf"] * 1.001).astype(np.float32))), os.path.join(outd, "gqi_small_odf.nii.gz"))
    assert chk.check(outd) == 1
    assert "NOT PINNED" in capsys.readouterr().out


def test_oracle_reproduces_rumba_fixture():
    """tests/golden/rumba_small.npz (tools/make_golden.py): the fixture the pinning kit hands to the reference's rumba_rec."""
    import rumba_oracle as R
    g = np.load(os.path.join(ROOT, "tests", "golden", "rumba_small.npz"))
    v, _ = O.load_sphere(362)
    r = R.rumba_rec(g["dwi"], g["mask"], g["bval"], g["bvec"], v, niter=int(g["niter"]), dtype=np.float64)
    assert np.array_equal(r["fodf"].astype(np.float32), g["fodf"]) and np.array_equal(r["peak_idx"], g["peak_idx"])
    assert r["snr_mean"] == float(g["snr_mean"]) and np.array_equal(r["gfa"], g["gfa"])
