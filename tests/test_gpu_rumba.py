"""GPU parity of rumba_rec (RUMBA-SD) through the C ABI against the numpy oracle (src/rusd.jl:266-340,
:419-636).  Tolerances: north_star's 1e-4 (amplitudes relative to the voxel's maximum, scalars relative), peak
indices exact except where the oracle's float64 amplitudes tie within 1e-5 of the voxel maximum."""
import numpy as np
import pytest

import fibers_oracle as O
import parity as P
import rumba_oracle as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import fibers_jl_b200 as F
    assert F.device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"
    return F


def _phantom(shape, seed, nb0=3, ndir=45, b=2000.0, fill=0.8, label_mask=False):
    from fibers_jl_b200 import phantom
    ph = phantom.gqi_phantom(shape, nb0=nb0, shells=((b, ndir),), seed=seed, mask_fill=fill)
    if label_mask:                                       # label map with a negative label: mask > 0 and mask != 0 differ
        m = ph["mask"].astype(np.int16) * 7
        m[0, 0, 0] = -3
        ph["mask"] = np.asfortranarray(m)
    return ph


def _compare(got, ref, what, nvert):
    mask_scale = np.abs(P.flat(ref["fodf"], nvert)).max(axis=1)
    err = P.odf_rel_err(got.fodf.vol, ref["fodf"])
    assert err < P.ODF_TOL, (what, err)
    out = {"fodf_max_rel_err": err}
    for name in ("fgm", "fcsf", "gfa", "var"):
        g = P.flat(getattr(got, name).vol).astype(np.float64); r = P.flat(ref[name]).astype(np.float64)
        assert np.all(g[r == 0] == 0), name
        d = np.abs(g - r) / np.maximum(np.abs(r), 1e-3 if name in ("fgm", "fcsf") else 1e-30)     # fractions: absolute 1e-7 floor
        out[name] = float(d.max())
        assert d.max() < P.SCALAR_TOL, (what, name, d.max())
    assert abs(got.snr_mean - ref["snr_mean"]) <= 1e-4 * abs(ref["snr_mean"]) and abs(got.snr_std - ref["snr_std"]) <= 1e-3 * abs(ref["snr_std"]) + 1e-6
    # peaks: same vertices in the same order, except near-ties of the float64 amplitudes
    gi = P.flat(got.peak_idx, 5).astype(np.int64); ri = P.flat(ref["peak_idx"], 5).astype(np.int64)
    bad = np.nonzero((gi != ri).any(axis=1))[0]
    f64 = P.flat(ref["fodf"], nvert).astype(np.float64)
    unexplained = 0
    for v in bad:
        tol = 1e-5 * f64[v].max()
        a = [i for i in gi[v] if i >= 0]; b = [i for i in ri[v] if i >= 0]
        # explained: the two lists hold the same amplitudes up to the tie tolerance (order / membership swaps among ties)
        va = np.sort(f64[v][a])[::-1]; vb = np.sort(f64[v][b])[::-1]
        unexplained += not (len(a) == len(b) and np.all(np.abs(va - vb) <= tol))
    assert unexplained == 0, (what, len(bad), unexplained)
    same = (gi == ri).all(axis=1)
    perr = 0.0
    for k in range(5):
        d = np.abs(P.flat(got.peak[k].vol, 3).astype(np.float64) - P.flat(ref["peak"][k], 3))[same]
        perr = max(perr, float(d.max()) if d.size else 0.0)
    assert perr < 1e-4
    out.update(peak_mismatch=int(len(bad)), peak_vec_abs_err=perr, voxels=int(mask_scale.size))
    print(f"[parity] {what}: " + ", ".join(f"{k}={v:.3g}" if isinstance(v, float) else f"{k}={v}" for k, v in out.items()))


@pytest.mark.parametrize("nsphere,niter,use_tv,kw", [
    (724, 40, True, {}),
    (724, 40, False, {}),
    (362, 25, True, {"ipat_factor": 2}),
    (642, 25, True, {"ncoils": 4, "coil_combine": "SoS-GRAPPA"}),
])
def test_rumba_parity(F, nsphere, niter, use_tv, kw):
    v, f = O.load_sphere(nsphere)
    ph = _phantom((10, 9, 7), seed=31 + nsphere, label_mask=(nsphere == 724 and use_tv))
    got = F.rumba_rec(F.MRI(ph["dwi"], ph["bval"], ph["bvec"]), F.MRI(ph["mask"]), F.ODF(v, f), niter=niter, use_tv=use_tv, **kw)
    ref = R.rumba_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, niter=niter, use_tv=use_tv, dtype=np.float64, **kw)
    _compare(got, ref, f"rumba sphere_{nsphere} niter={niter} tv={use_tv} {kw}", nsphere // 2)
    assert (P.flat(got.peak_idx, 5)[:, 0] >= 0).mean() > 0.3


def test_rumba_long_run_and_multishell(F):
    """300 iterations on a multi-shell protocol (minimum b = 5, not 0): the fp32 iteration stays within tolerance of float64."""
    v, f = O.load_sphere(724)
    from fibers_jl_b200 import phantom
    ph = phantom.gqi_phantom((7, 6, 5), nb0=4, shells=((1000.0, 24), (3000.0, 32)), seed=77, mask_fill=0.9)
    ph["bval"] = ph["bval"].copy(); ph["bval"][:4] = 5.0
    got = F.rumba_rec(F.MRI(ph["dwi"], ph["bval"], ph["bvec"]), F.MRI(ph["mask"]), niter=300)
    ref = R.rumba_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, niter=300, dtype=np.float64)
    _compare(got, ref, "rumba 300 iterations, 2 shells", 362)


def test_rumba_edge_cases_and_errors(F):
    v, f = O.load_sphere(724)
    ph = _phantom((6, 5, 4), seed=9)
    dwi = F.MRI(ph["dwi"], ph["bval"], ph["bvec"])
    # empty mask: everything zero, no iteration runs
    z = F.rumba_rec(dwi, F.MRI(np.zeros((6, 5, 4), np.uint8)), niter=5)
    assert np.all(z.fodf.vol == 0) and np.all(z.peak_idx == -1) and z.snr_mean == 0
    # niter = 0: start value only (uniform fodf), as the reference's loop `for iter in 1:0`
    r0 = F.rumba_rec(dwi, F.MRI(ph["mask"]), niter=0)
    ref0 = R.rumba_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, niter=0, dtype=np.float64)
    assert P.odf_rel_err(r0.fodf.vol, ref0["fodf"]) < 1e-5
    with pytest.raises(RuntimeError, match="Unknown coil combine mode"):
        F.rumba_rec(dwi, F.MRI(ph["mask"]), coil_combine="SENSE")
    with pytest.raises(RuntimeError, match="iPAT factor must be a positive integer"):
        F.rumba_rec(dwi, F.MRI(ph["mask"]), ipat_factor=0)
    with pytest.raises(RuntimeError, match="Missing gradient table"):
        F.rumba_rec(F.MRI(ph["dwi"], ph["bval"]), F.MRI(ph["mask"]))
