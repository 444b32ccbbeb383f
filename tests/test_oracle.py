"""CPU tests of the oracle itself: known answers, invariants, fp32-vs-fp64, numpy-vs-C port,
and the committed golden fixtures (tools/make_golden.py)."""
import os

import numpy as np
import pytest

import fibers_oracle as O
import parity as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _phantom():
    from fibers_jl_b200 import phantom
    return phantom


# ------------------------------------------------------------------ sphere tables (src/odf.jl)
@pytest.mark.parametrize("n,nface,degs", [(362, 720, {5: 6, 6: 175}), (642, 1280, {5: 6, 6: 315}), (724, 1444, None)])
def test_sphere_tables(n, nface, degs):
    v, f = O.load_sphere(n)
    M = n // 2
    assert v.shape == (n, 3) and v.dtype == np.float32 and f.shape == (nface, 3)
    assert np.array_equal(v[M:], -v[:M]), "antipodal symmetry must be exact in fp32"
    assert np.abs(np.linalg.norm(v.astype(np.float64), axis=1) - 1).max() < 2e-3
    assert f.min() == 1 and f.max() == n
    nv, ne = n, len({tuple(sorted(e)) for t in f for e in ((t[0], t[1]), (t[1], t[2]), (t[0], t[2]))})
    assert nv - ne + nface == 2, "Euler characteristic of a sphere"
    nbr = O.neighbour_table(O.fold_faces(f, M), M)
    deg = (nbr >= 0).sum(axis=1)
    assert deg.max() <= 7 and deg.min() >= 5
    if degs:
        assert {int(k): int(c) for k, c in zip(*np.unique(deg, return_counts=True))} == degs
    assert not (nbr == np.arange(M)[:, None]).any(), "no degenerate folded faces"


# ------------------------------------------------------------------ find_peaks!
def test_find_peaks_kat_and_equivalence(sphere642):
    v, f = sphere642
    M = 321
    ff = O.fold_faces(f, M); nbr = O.neighbour_table(ff, M)
    z = np.load(os.path.join(GOLD, "peaks_kat.npz"))
    odf, exp = z["odf"], z["expected"]
    assert list(exp[0]) == [-1, -1, -1, 0]          # plateau
    assert list(exp[1]) == [10, -1, -1, 1]          # single peak
    assert list(exp[2]) == [-1, -1, -1, 0]          # adjacent tie kills both
    assert list(exp[3]) == [-1, -1, -1, 0]          # negative maximum never counts
    assert list(exp[4]) == [3, 200, 310, 3]         # equal values keep ascending index (stable sort)
    assert list(exp[5]) == [100, 250, 150, 4]       # top 3 by value of 4 peaks
    idx, nvalid = O.find_peaks_batch(odf, nbr)
    for i, o in enumerate(odf):
        isort, nv = O.find_peaks_literal(o, ff)
        want = [isort[k] if k < min(nv, 3) else -1 for k in range(3)]
        assert want + [nv] == list(exp[i])
        assert list(idx[i]) == want and nvalid[i] == nv


@pytest.mark.parametrize("n", [362, 724])
def test_find_peaks_equivalence_other_spheres(n):
    v, f = O.load_sphere(n)
    M = n // 2
    ff = O.fold_faces(f, M); nbr = O.neighbour_table(ff, M)
    rng = np.random.default_rng(n)
    odf = (np.round(rng.normal(size=(60, M)) * 3) / 2).astype(np.float32)
    idx, nvalid = O.find_peaks_batch(odf, nbr)
    for i, o in enumerate(odf):
        isort, nv = O.find_peaks_literal(o, ff)
        assert list(idx[i]) == [isort[k] if k < min(nv, 3) else -1 for k in range(3)] and nvalid[i] == nv


# ------------------------------------------------------------------ eigen / DTI
def test_eig3_against_lapack():
    rng = np.random.default_rng(0)
    A = rng.normal(size=(5000, 3, 3)) * 1e-3
    A = (A + A.transpose(0, 2, 1)).astype(np.float32)
    for T, tol in ((np.float32, 2e-5), (np.float64, 1e-12)):
        a = A.astype(T)
        w, V = O.eig3_sym(a[:, 0, 0], a[:, 1, 0], a[:, 2, 0], a[:, 1, 1], a[:, 2, 1], a[:, 2, 2], T)
        w0, V0 = np.linalg.eigh(A.astype(np.float64))
        assert np.all(np.diff(w, axis=1) >= 0)
        assert (np.abs(w - w0) / np.abs(w0).max(axis=1, keepdims=True)).max() < tol
        res = np.einsum("nij,njk->nik", A.astype(np.float64), V.astype(np.float64)) - V * w[:, None, :]
        assert np.abs(res).max() / 1e-3 < (1e-4 if T == np.float32 else 1e-11)
    # diagonal branch
    w, V = O.eig3_sym([3.0, 1.0], [0, 0], [0, 0], [1.0, 2.0], [0, 0], [2.0, 3.0], np.float32)
    assert np.array_equal(w, [[1, 2, 3], [1, 2, 3]])
    assert np.array_equal(np.abs(V[0]), [[0, 0, 1], [1, 0, 0], [0, 1, 0]])


def test_dti_noise_free_kat():
    ph = _phantom().dti_phantom((12, 12, 6), snr=0, inject=False, seed=3)
    mask = np.ones((12, 12, 6), np.uint8)
    r = O.dti_fit(ph["dwi"], mask, ph["bval"], ph["bvec"], np.float64)
    assert np.abs(P.flat(r["eigval1"]) - ph["l1"]).max() / 1e-3 < 1e-5
    assert np.abs(P.flat(r["eigval3"]) - ph["l3"]).max() / 1e-3 < 1e-5
    assert np.abs((P.flat(r["eigvec1"], 3) * ph["e1"]).sum(axis=1)).min() > 1 - 1e-8
    assert np.abs(P.flat(r["s0"]) - ph["S0"]).max() / 1500 < 1e-5
    r32 = O.dti_fit(ph["dwi"], mask, ph["bval"], ph["bvec"], np.float32)
    assert np.abs(r32["fa"] - r["fa"]).max() < 1e-4


def test_dti_branch_rules_and_maps():
    ph = _phantom().dti_phantom((12, 10, 6), seed=4)
    r = O.dti_fit(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], np.float32)
    assert set(np.unique(r["kind"])) == {0, 1, 2}
    S = P.flat(ph["dwi"], 31); m = P.flat(ph["mask"]) != 0
    npos = (S > 0).sum(axis=1)
    expect = m & ((npos == 31) | ((npos > 6) & (S[:, 0] > 0)))        # src/dti.jl:294-303
    assert np.array_equal(P.flat(r["valid"]), expect)
    assert np.all(P.flat(r["fa"])[~expect] == 0)
    rd, md, fa = O.dti_maps(np.float32(3), np.float32(2), np.float32(1))
    assert rd == 1.5 and md == 2.0 and abs(fa - np.sqrt(1.5 * 2 / 14)) < 1e-7
    with np.errstate(all="ignore"):
        assert np.isnan(O.dti_maps(np.float32(0), np.float32(0), np.float32(0))[2])    # no clamp: 0/0


# ------------------------------------------------------------------ GQI / DSI
def test_gqi_matrix_properties(sphere642):
    v, f = sphere642
    bval, bvec = _phantom().shells_table(18, [(1000.0, 90), (2000.0, 90), (3000.0, 90)])
    A = O.gqi_matrix(bval, bvec, v)
    assert A.shape == (321, 288) and A.dtype == np.float32
    assert np.all(A[:, :18] == 1), "b = 0 columns are sinc(0) = 1"
    assert A.min() > -0.2173 and A.max() <= 1
    assert np.abs(A - O.gqi_matrix(bval, bvec, v, dtype=np.float64)).max() < 1e-6


def test_gqi_single_fibre_peak_and_invariants(sphere642):
    v, f = sphere642
    ph = _phantom()
    bval, bvec = ph.shells_table(2, [(1000.0, 60), (3000.0, 60)])
    rng = np.random.default_rng(1)
    e = rng.normal(size=(40, 3)); e /= np.linalg.norm(e, axis=1, keepdims=True)
    S = 1000 * np.exp(-bval[None] * (2e-4 + 1.5e-3 * (e @ bvec.T.astype(np.float64)) ** 2))
    dwi = np.asfortranarray(S.reshape(40, 1, 1, -1).astype(np.float32))
    r = O.gqi_rec(dwi, np.ones((40, 1, 1), np.uint8), bval, bvec, v, f, 1.25, np.float64)
    idx = r["peak_idx"][:, 0, 0, 0]
    best = np.argmax(np.abs(v[:321].astype(np.float64) @ e.T), axis=0)
    ang = np.abs((v[idx] * e).sum(axis=1))
    assert np.all(ang > np.cos(np.deg2rad(8))), "peak within one mesh spacing of the fibre"
    assert (idx == best).mean() > 0.5
    assert np.all(r["nvalid"] >= 1) and np.all(r["qa"][0] >= 0) and np.all(r["qa"][0] >= r["qa"][1])
    assert np.all((r["peak_idx"] >= -1) & (r["peak_idx"] < 321))


def test_dsi_matrix_form_equals_fft_form(sphere642):
    v, f = sphere642
    ph = _phantom().dsi_phantom((6, 5, 4), seed=3)
    d64 = O.dsi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 32, np.float64)
    Mo, Mp, cvol, dscale = O.dsi_matrices(ph["bval"], ph["bvec"], v, 32)
    assert cvol == 0 and dscale == 4096.0
    S = np.maximum(P.flat(ph["dwi"], 515).astype(np.float64), 0)
    c = P.flat(d64["computed"])
    den = dscale * S[c, cvol]
    odf = (S[c] @ Mo.T) / den[:, None]; pdf = (S[c] @ Mp.T) / den[:, None]
    assert np.abs(odf - P.flat(d64["odf"], 321)[c]).max() / np.abs(odf).max() < 1e-12
    assert np.abs(pdf - P.flat(d64["pdf"], 515)[c]).max() / np.abs(pdf).max() < 1e-12
    d32 = O.dsi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 32, np.float32)
    assert P.odf_rel_err(d32["odf"], d64["odf"]) < 1e-5 and P.odf_rel_err(d32["pdf"], d64["pdf"]) < 1e-5
    # duplicate q-points: last write wins (a second b0 volume appended at the end)
    bval = np.concatenate([ph["bval"], [0]]).astype(np.float32)
    bvec = np.concatenate([ph["bvec"], [[0, 0, 0]]]).astype(np.float32)
    dwi = np.asfortranarray(np.concatenate([ph["dwi"], ph["dwi"][..., :1] * 1.1], axis=3))
    d2 = O.dsi_rec(dwi, ph["mask"], bval, bvec, v, f, 32, np.float64)
    Mo2, Mp2, cvol2, _ = O.dsi_matrices(bval, bvec, v, 32)
    assert cvol2 == 515 and np.all(Mo2[:, 0] == 0) and np.array_equal(Mp2[0], Mp2[515])
    S2 = np.maximum(P.flat(dwi, 516).astype(np.float64), 0)
    c2 = P.flat(d2["computed"])
    odf2 = (S2[c2] @ Mo2.T) / (4096.0 * S2[c2, 515])[:, None]
    assert np.abs(odf2 - P.flat(d2["odf"], 321)[c2]).max() / np.abs(odf2).max() < 1e-12


# ------------------------------------------------------------------ numpy oracle vs C port vs golden
def test_c_port_matches_numpy_oracle(sphere642):
    import c_oracle as CO
    v, f = sphere642
    ph = _phantom().gqi_phantom((8, 7, 5), seed=12, mask_fill=0.7)
    c = CO.gqi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, nthreads=2)
    r = O.gqi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 1.25, np.float32)
    assert P.odf_rel_err(c["odf"], r["odf"]) < 2e-6
    assert np.array_equal(c["peak_idx"], r["peak_idx"])
    assert np.nanmax(np.abs(c["qa"][0] - r["qa"][0])) < 1e-5
    pd = _phantom().dsi_phantom((5, 4, 3), seed=13)
    c = CO.dsi_rec(pd["dwi"], pd["mask"], pd["bval"], pd["bvec"], v, f)
    r = O.dsi_rec(pd["dwi"], pd["mask"], pd["bval"], pd["bvec"], v, f, 32, np.float32)
    assert P.odf_rel_err(c["odf"], r["odf"]) < 2e-5 and P.odf_rel_err(c["pdf"], r["pdf"]) < 2e-5
    assert (c["peak_idx"] != r["peak_idx"]).any(axis=-1).mean() < 0.02
    pt = _phantom().dti_phantom((12, 10, 6), seed=14)
    c = CO.dti_fit(pt["dwi"], pt["mask"], pt["bval"], pt["bvec"], nthreads=2)
    r = O.dti_fit(pt["dwi"], pt["mask"], pt["bval"], pt["bvec"], np.float32)
    assert np.array_equal(c["valid"], r["valid"])
    assert np.abs(c["fa"] - r["fa"]).max() < 1e-4 and np.abs(c["md"] - r["md"]).max() / 1e-3 < 1e-4


def test_golden_fixtures_reproduce(sphere642):
    v, f = sphere642
    g = np.load(os.path.join(GOLD, "gqi_small.npz"))
    for T, tol in ((np.float64, 1e-6), (np.float32, 1e-5)):
        r = O.gqi_rec(g["dwi"], g["mask"], g["bval"], g["bvec"], v, f, 1.25, T)
        assert P.odf_rel_err(r["odf"], g["odf"]) < tol
        assert np.array_equal(r["peak_idx"], g["peak_idx"])
        assert np.array_equal(r["computed"], g["computed"])
    d = np.load(os.path.join(GOLD, "dsi_small.npz"))
    r = O.dsi_rec(d["dwi"], d["mask"], d["bval"], d["bvec"], v, f, 32, np.float64)
    assert P.odf_rel_err(r["odf"], d["odf"]) < 1e-6 and P.odf_rel_err(r["pdf"], d["pdf"]) < 1e-6
    assert np.array_equal(r["peak_idx"], d["peak_idx"])
    t = np.load(os.path.join(GOLD, "dti_small.npz"))
    r = O.dti_fit(t["dwi"], t["mask"], t["bval"], t["bvec"], np.float64)
    assert np.array_equal(r["valid"], t["valid"]) and np.array_equal(r["kind"], t["kind"])
    assert np.allclose(r["fa"], t["fa"], rtol=1e-9, atol=1e-12, equal_nan=True)
    assert np.allclose(r["eigval1"], t["eigval1"], rtol=1e-9, atol=1e-15)
