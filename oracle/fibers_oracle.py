"""CPU oracle for the Fibers.jl voxel-wise diffusion reconstruction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (`fibers.jl_b200/`, the
C-ABI library) may import or call this module; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs use it, and there only as the checker or the timed CPU baseline.

PARITY UNPINNED: the reference (lincbrain/Fibers.jl v1.0.0, pure Julia) ships an
empty test suite (test/runtests.jl:4-6), no golden vectors, and Julia is not
installed in this image, so the reference cannot be executed here.  This file
restates the reference algorithm line by line in numpy; the third-party pieces
the reference calls and that are NOT under /root/reference are restated from
their published algorithms:
  * StaticArrays.jl (compat 1.4.4) `eigen(Symmetric(::SMatrix{3,3}))`
    (call site src/dti.jl:311): closed-form trigonometric eigenvalues,
    eigenvectors from the best-conditioned row cross product followed by a
    projected 2x2 problem (Eberly / Kopp style).  -> `eig3_sym`
  * LinearAlgebra.pinv (LAPACK SVD)  (src/dti.jl:72,143,207,298) -> numpy pinv
  * FFTW.jl 16^3 complex FFT (src/dsi.jl:91,219) -> numpy.fft
  * Interpolations.jl BSpline(Linear()) (src/dsi.jl:230-238) -> `_trilinear`
  * Base.sinc, Base.sortperm!(rev=true) (stable), Statistics.mean.

Every function exists in two precisions selected by `dtype`:
  np.float32 : same operation order as the reference, explicit fp32 arithmetic
               (numpy never contracts to FMA) -- "what Julia would print".
  np.float64 : the same formulas in double -- "truth", used to decide whether a
               GPU/fp32 difference is a genuine error or a near-tie.

All volumes are Julia column-major `[nx, ny, nz, nframes]`; numpy arrays passed
in/out use the same index order (use order='F' arrays for speed).
"""
from __future__ import annotations

import numpy as np

NPEAK = 3  # src/gqi.jl:119, src/dsi.jl:181


# --------------------------------------------------------------------------
# Sphere tables (src/odf.jl) -- loaded from the exported fixture
# --------------------------------------------------------------------------
def load_sphere(n: int = 642, path: str | None = None):
    """Return (vertices float32 [2M,3], faces int32 [F,3] 1-based) of sphere_<n>.

    reference: src/odf.jl:14 (362), :1104 (642, default), :3034 (724)."""
    import os
    if path is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                            "fibers.jl_b200", "data", "spheres.npz")
    z = np.load(path)
    return z[f"vertices_{n}"], z[f"faces_{n}"]


def fold_faces(faces: np.ndarray, nvert: int) -> np.ndarray:
    """faces[faces .> nvert] .-= nvert  (src/gqi.jl:63-64, src/dsi.jl:121-122)."""
    f = np.array(faces, dtype=np.int64, copy=True)
    f[f > nvert] -= nvert
    return f


def neighbour_table(faces_folded: np.ndarray, nvert: int, width: int = 8) -> np.ndarray:
    """[nvert, width] 0-based neighbour indices of the folded mesh, padded with -1.

    Equivalent restatement of the face-based suppression in find_peaks!
    (src/gqi.jl:185-196): a vertex survives iff it is strictly greater than every
    vertex it shares a folded face with."""
    nb = [set() for _ in range(nvert)]
    for a, b, c in (faces_folded - 1):
        nb[a].update((b, c)); nb[b].update((a, c)); nb[c].update((a, b))
    tab = -np.ones((nvert, width), dtype=np.int32)
    for i, s in enumerate(nb):
        s = sorted(s)  # NB: a vertex listed with itself in a face would kill itself (>=)
        assert len(s) <= width, (i, len(s))
        tab[i, :len(s)] = s
    return tab


# --------------------------------------------------------------------------
# find_peaks!  (src/gqi.jl:180-201)
# --------------------------------------------------------------------------
def find_peaks_literal(o: np.ndarray, faces_folded: np.ndarray):
    """Literal single-voxel restatement of find_peaks! (src/gqi.jl:180-201).

    Returns (isort 0-based int array [nvert], nvalid)."""
    f = faces_folded - 1
    odf_peak = o.copy()
    o1, o2, o3 = o[f[:, 0]], o[f[:, 1]], o[f[:, 2]]
    odf_peak[f[(o2 >= o1) | (o3 >= o1), 0]] = 0      # :185-188
    odf_peak[f[(o1 >= o2) | (o3 >= o2), 1]] = 0      # :189-192
    odf_peak[f[(o2 >= o3) | (o1 >= o3), 2]] = 0      # :193-196
    # sortperm!(rev=true) is stable: equal keys keep ascending index (:198)
    isort = np.argsort(-odf_peak, kind="stable")
    nvalid = int(np.count_nonzero(odf_peak > 0))      # :200
    return isort, nvalid


def find_peaks_batch(odf: np.ndarray, nbr: np.ndarray, npeak: int = NPEAK):
    """Vectorised equivalent over voxels: odf [nv, M] -> (idx [nv, npeak] 0-based
    or -1, nvalid [nv]).  Candidates {v: o[v] > 0 and o[v] > o[n] for all n in N(v)}
    ranked by (value desc, index asc)  (verified against
    find_peaks_literal in tests/test_oracle.py)."""
    nv, M = odf.shape
    pad = np.concatenate([odf, np.full((nv, 1), -np.inf, odf.dtype)], axis=1)
    alive = odf > 0
    for k in range(nbr.shape[1]):
        col = nbr[:, k].astype(np.int64)
        col = np.where(col < 0, M, col)
        alive &= odf > pad[:, col]
    key = np.where(alive, odf, -np.inf)
    order = np.argsort(-key, axis=1, kind="stable")[:, :npeak]
    nvalid = alive.sum(axis=1)
    idx = np.where(np.arange(npeak)[None, :] < np.minimum(nvalid, npeak)[:, None], order, -1)
    return idx.astype(np.int32), nvalid.astype(np.int32)


# --------------------------------------------------------------------------
# 3x3 symmetric eigen-decomposition (StaticArrays.jl closed form; src/dti.jl:311)
# --------------------------------------------------------------------------
def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def eig3_sym(a11, a12, a13, a22, a23, a33, dtype=np.float32):
    """Eigen-decomposition of symmetric 3x3 matrices (vectorised over voxels).

    Restates StaticArrays.jl `_eig(::Size{(3,3)}, ::RealHermSymComplexHerm)` as
    called by `eigen(Symmetric(D, :L))` at src/dti.jl:311.  Returns
    (vals [n,3] ascending, vecs [n,3,3] with vecs[:, :, k] the k-th eigenvector).
    Eigenvector signs are whatever the cross products give (arbitrary)."""
    T = dtype
    a11, a12, a13, a22, a23, a33 = (np.asarray(x, dtype=T).reshape(-1) for x in
                                    (a11, a12, a13, a22, a23, a33))
    n = a11.shape[0]
    vals = np.zeros((n, 3), T)
    vecs = np.zeros((n, 3, 3), T)
    with np.errstate(all="ignore"):
        p1 = a12 * a12 + a13 * a13 + a23 * a23
        diag = p1 == 0
        q = (a11 + a22 + a33) / T(3)
        p2 = (a11 - q) ** 2 + (a22 - q) ** 2 + (a33 - q) ** 2 + T(2) * p1
        p = np.sqrt(p2 / T(6))
        invp = T(1) / p
        b11 = (a11 - q) * invp; b22 = (a22 - q) * invp; b33 = (a33 - q) * invp
        b12 = a12 * invp; b13 = a13 * invp; b23 = a23 * invp
        # det(B) = x0 . (x1 x x2) with columns x0=(b11,b12,b13) x1=(b12,b22,b23) x2=(b13,b23,b33)
        c = _cross((b12, b22, b23), (b13, b23, b33))
        r = (b11 * c[0] + b12 * c[1] + b13 * c[2]) / T(2)
        pi = T(np.pi)
        phi = np.where(r <= -1, pi / T(3), np.where(r >= 1, T(0), np.arccos(np.clip(r, -1, 1)) / T(3))).astype(T)
        eig3 = q + T(2) * p * np.cos(phi)
        eig1 = q + T(2) * p * np.cos(phi + (T(2) * pi / T(3)))
        eig2 = T(3) * q - eig1 - eig3
        swap = r > 0
        e1 = np.where(swap, eig3, eig1)      # eigenvalue used for the first eigenvector
        e3 = np.where(swap, eig1, eig3)
        # first eigenvector: best of the three row cross products of A - e1*I
        r1 = (a11 - e1, a12, a13); r2 = (a12, a22 - e1, a23); r3 = (a13, a23, a33 - e1)
        n1 = r1[0] ** 2 + r1[1] ** 2 + r1[2] ** 2
        n2 = r2[0] ** 2 + r2[1] ** 2 + r2[2] ** 2
        n3 = r3[0] ** 2 + r3[1] ** 2 + r3[2] ** 2
        r12 = _cross(r1, r2); r23 = _cross(r2, r3); r31 = _cross(r3, r1)
        n12 = r12[0] ** 2 + r12[1] ** 2 + r12[2] ** 2
        n23 = r23[0] ** 2 + r23[1] ** 2 + r23[2] ** 2
        n31 = r31[0] ** 2 + r31[1] ** 2 + r31[2] ** 2
        cA = n12 * n3 > n23 * n1
        use12 = cA & (n12 * n3 > n31 * n2)
        use23 = (~cA) & (n23 * n1 > n31 * n2)
        sel = lambda x12, x23, x31: np.where(use12, x12, np.where(use23, x23, x31))
        nn = sel(n12, n23, n31)
        sn = np.sqrt(nn)
        v1 = tuple(sel(r12[k], r23[k], r31[k]) / sn for k in range(3))
        # orthonormal complement
        cB = np.abs(v1[0]) < np.abs(v1[1])
        dB1 = np.sqrt(v1[0] ** 2 + v1[2] ** 2)
        dB2 = np.sqrt(v1[1] ** 2 + v1[2] ** 2)
        z = np.zeros_like(a11)
        o1 = (np.where(cB, -v1[2] / dB1, z), np.where(cB, z, v1[2] / dB2),
              np.where(cB, v1[0] / dB1, -v1[1] / dB2))
        o2 = _cross(v1, o1)
        # projected 2x2 problem of A - eig2*I on {o1, o2}
        Ao1 = (a11 * o1[0] + a12 * o1[1] + a13 * o1[2],
               a12 * o1[0] + a22 * o1[1] + a23 * o1[2],
               a13 * o1[0] + a23 * o1[1] + a33 * o1[2])
        Ao2 = (a11 * o2[0] + a12 * o2[1] + a13 * o2[2],
               a12 * o2[0] + a22 * o2[1] + a23 * o2[2],
               a13 * o2[0] + a23 * o2[1] + a33 * o2[2])
        c11 = o1[0] * Ao1[0] + o1[1] * Ao1[1] + o1[2] * Ao1[2] - eig2
        c12 = o1[0] * Ao2[0] + o1[1] * Ao2[1] + o1[2] * Ao2[2]
        c22 = o2[0] * Ao2[0] + o2[1] * Ao2[1] + o2[2] * Ao2[2] - eig2
        s11 = c11 * c11; s12 = c12 * c12; s22 = c22 * c22
        # four branches + degenerate
        brA = s11 >= s22
        brA1 = brA & (s11 >= s12)          # tmp = c12/c11 ; pp2 = 1/sqrt(1+tmp^2); pp1 = tmp*pp2
        brA2 = brA & ~(s11 >= s12)         # tmp = c11/c12 ; pp1 = 1/sqrt(1+tmp^2); pp2 = tmp*pp1
        brB1 = (~brA) & (s22 >= s12)       # tmp = c12/c22 ; pp1 = ...; pp2 = tmp*pp1
        brB2 = (~brA) & ~(s22 >= s12)      # tmp = c22/c12 ; pp2 = ...; pp1 = tmp*pp2
        degen = brA & ~((s11 > 0) | (s12 > 0))
        tmp = np.where(brA1, c12 / c11, np.where(brA2, c11 / c12, np.where(brB1, c12 / c22, c22 / c12)))
        base = T(1) / np.sqrt(T(1) + tmp * tmp)
        oth = tmp * base
        pp1 = np.where(brA1 | brB2, oth, base)
        pp2 = np.where(brA1 | brB2, base, oth)
        v2 = tuple(np.where(degen, o1[k], pp1 * o1[k] - pp2 * o2[k]) for k in range(3))
        v3 = _cross(v1, v2)
        # un-swap
        V1 = tuple(np.where(swap, v3[k], v1[k]) for k in range(3))
        V3 = tuple(np.where(swap, v1[k], v3[k]) for k in range(3))
        vals[:, 0] = np.where(swap, e3, e1); vals[:, 1] = eig2; vals[:, 2] = np.where(swap, e1, e3)
        for k in range(3):
            vecs[:, k, 0] = V1[k]; vecs[:, k, 1] = v2[k]; vecs[:, k, 2] = V3[k]
    if diag.any():
        # diagonal matrix: sorted diagonal, unit vectors (StaticArrays p1 == 0 branch)
        idx = np.nonzero(diag)[0]
        d = np.stack([a11[idx], a22[idx], a33[idx]], axis=1)
        for j, i in enumerate(idx):
            x, y, zz = d[j]
            if x < y:
                order = (0, 1, 2) if y < zz else ((2, 0, 1) if zz < x else (0, 2, 1))
            else:
                order = (1, 0, 2) if x < zz else ((2, 1, 0) if zz < y else (1, 2, 0))
            vals[i] = d[j, list(order)]
            vecs[i] = 0
            for k, oi in enumerate(order):
                vecs[i, oi, k] = 1
    return vals, vecs


# --------------------------------------------------------------------------
# DTI / ADC   (src/dti.jl)
# --------------------------------------------------------------------------
def dti_design(bval, bvec, dtype=np.float32):
    """DTIwork: ib0, A [N,7], pA [7,N]   (src/dti.jl:110-155)."""
    T = dtype
    bval = np.asarray(bval, T); bvec = np.asarray(bvec, T)
    ib0 = bval == bval.min()                                     # :117
    A = np.empty((bval.shape[0], 7), T)
    A[:, 0] = bvec[:, 0] ** 2                                    # :133-138
    A[:, 1] = T(2) * bvec[:, 0] * bvec[:, 1]
    A[:, 2] = T(2) * bvec[:, 0] * bvec[:, 2]
    A[:, 3] = bvec[:, 1] ** 2
    A[:, 4] = T(2) * bvec[:, 1] * bvec[:, 2]
    A[:, 5] = bvec[:, 2] ** 2
    A[:, :6] *= -bval[:, None]                                   # :140
    A[:, 6] = 1                                                  # :142
    pA = np.linalg.pinv(A).astype(T)                             # :145
    return ib0, A, pA


def adc_design(bval, dtype=np.float32):
    """ADCwork (src/dti.jl:49-83)."""
    T = dtype
    bval = np.asarray(bval, T)
    ib0 = bval == bval.min()
    A = np.stack([-bval, np.ones_like(bval)], axis=1).astype(T)
    return ib0, A, np.linalg.pinv(A).astype(T)


def dti_maps(l1, l2, l3, dtype=np.float32):
    """rd, md, fa   (src/dti.jl:325-335); no clamp, 0/0 -> NaN."""
    T = dtype
    with np.errstate(all="ignore"):
        rd = l2 + l3
        md = (l1 + rd) / T(3)
        rd = rd / T(2)
        fa = np.sqrt(((l1 - md) ** 2 + (l2 - md) ** 2 + (l3 - md) ** 2) /
                     (l1 ** 2 + l2 ** 2 + l3 ** 2) * T(1.5))
    return rd, md, fa


def _solve_linear_fit(S, ib0, A, pA, T):
    """Shared branch logic of adc_fit / dti_fit_ls voxel fits.

    S [nv, N] (already T).  Returns (d [nv, ncol], valid [nv] bool, kind [nv] int8
    0 = zero output, 1 = full path, 2 = partial (per-voxel pinv) path)."""
    nv, N = S.shape
    ipos = S > 0                                                 # dti.jl:291
    npos = ipos.sum(axis=1)
    full = npos == N
    part = (~full) & (npos > 6) & (ipos[:, ib0].any(axis=1))     # dti.jl:297 / :206
    d = np.zeros((nv, A.shape[1]), T)
    if full.any():
        with np.errstate(all="ignore"):
            d[full] = (np.log(S[full]) @ pA.T).astype(T)         # :294-296 (sgemv)
    for i in np.nonzero(part)[0]:                                # :298 per-voxel pinv
        m = ipos[i]
        d[i] = (np.linalg.pinv(A[m]).astype(T) @ np.log(S[i, m])).astype(T)
    kind = np.zeros(nv, np.int8); kind[full] = 1; kind[part] = 2
    return d, full | part, kind


def adc_fit(dwi, mask, bval, dtype=np.float32):
    """adc_fit(dwi::MRI, mask::MRI) -> (adc, s0)   (src/dti.jl:164-213)."""
    T = dtype
    shp = dwi.shape[:3]
    ib0, A, pA = adc_design(bval, T)
    m = (np.asarray(mask).reshape(shp) != 0).reshape(-1, order="F")
    S = np.asarray(dwi, T).reshape((-1, dwi.shape[3]), order="F")[m]
    d, valid, _ = _solve_linear_fit(S, ib0, A, pA, T)
    adc = np.zeros(m.shape[0], T); s0 = np.zeros(m.shape[0], T)
    a = np.where(valid, d[:, 0], 0).astype(T)
    with np.errstate(all="ignore"):
        e = np.where(valid, np.exp(d[:, 1]), 0).astype(T)
    adc[m] = a; s0[m] = e
    return adc.reshape(shp, order="F"), s0.reshape(shp, order="F")


def dti_fit(dwi, mask, bval, bvec, dtype=np.float32, eig="closed"):
    """dti_fit(dwi::MRI, mask::MRI) -> DTI   (src/dti.jl:221-316).

    Returns dict with s0, eigval1..3, rd, md, fa [nx,ny,nz] and eigvec1..3
    [nx,ny,nz,3], plus `valid` (bool: voxel took the full or partial path) and
    `kind`.  eig='closed' follows StaticArrays' closed form in `dtype`;
    eig='lapack' uses numpy eigh in float64 (truth for the decomposition)."""
    T = dtype
    shp = dwi.shape[:3]
    nvox = int(np.prod(shp))
    ib0, A, pA = dti_design(bval, bvec, T)
    m = (np.asarray(mask).reshape(shp) != 0).reshape(-1, order="F")
    S = np.asarray(dwi, T).reshape((-1, dwi.shape[3]), order="F")[m]
    d, valid, kind = _solve_linear_fit(S, ib0, A, pA, T)
    with np.errstate(all="ignore"):
        s0 = np.exp(d[:, 6]).astype(T)                          # :305
    if eig == "closed":
        vals, vecs = eig3_sym(d[:, 0], d[:, 1], d[:, 2], d[:, 3], d[:, 4], d[:, 5], T)
    else:
        D = np.zeros((d.shape[0], 3, 3), np.float64)
        D[:, 0, 0] = d[:, 0]; D[:, 1, 0] = D[:, 0, 1] = d[:, 1]; D[:, 2, 0] = D[:, 0, 2] = d[:, 2]
        D[:, 1, 1] = d[:, 3]; D[:, 2, 1] = D[:, 1, 2] = d[:, 4]; D[:, 2, 2] = d[:, 5]
        vals, vecs = np.linalg.eigh(D)
        vals = vals.astype(T); vecs = vecs.astype(T)
    l1, l2, l3 = vals[:, 2], vals[:, 1], vals[:, 0]             # :313
    rd, md, fa = dti_maps(l1, l2, l3, T)
    out = {}

    def put(name, x, nfr=1):
        full = np.zeros((nvox, nfr), T)
        xx = np.where(valid[:, None], np.asarray(x, T).reshape(-1, nfr), 0)
        full[m] = xx
        out[name] = full.reshape(shp + ((nfr,) if nfr > 1 else ()), order="F")

    put("s0", s0); put("eigval1", l1); put("eigval2", l2); put("eigval3", l3)
    put("eigvec1", vecs[:, :, 2], 3); put("eigvec2", vecs[:, :, 1], 3); put("eigvec3", vecs[:, :, 0], 3)
    put("rd", rd); put("md", md); put("fa", fa)
    v = np.zeros(nvox, bool); v[m] = valid
    k = np.zeros(nvox, np.int8); k[m] = kind
    out["valid"] = v.reshape(shp, order="F"); out["kind"] = k.reshape(shp, order="F")
    return out


# --------------------------------------------------------------------------
# GQI   (src/gqi.jl)
# --------------------------------------------------------------------------
def _sinc(x):
    """Base.sinc: sin(pi x)/(pi x), sinc(0) = 1 (computed in x's dtype)."""
    T = x.dtype.type
    px = T(np.pi) * x
    with np.errstate(all="ignore"):
        return np.where(x == 0, T(1), np.sin(px) / px).astype(T)


def gqi_matrix(bval, bvec, vertices, sigma=1.25, dtype=np.float32):
    """GQIwork system matrix A [M, N]   (src/gqi.jl:66-69).  fp32 constant chain:
    bq = bvec .* (sqrt.(bval*T(0.01506)) * T(sigma/pi));  A = sinc.(V[M+1:end,:]*bq')."""
    T = dtype
    bval = np.asarray(bval, T); bvec = np.asarray(bvec, T)
    M = vertices.shape[0] // 2
    if T == np.float32:
        sig = np.float32(np.float32(sigma) / np.float32(np.pi))   # Float32/Irrational -> Float32
    else:
        sig = T(np.float64(np.float32(sigma)) / np.pi)
    bq = bvec * (np.sqrt(bval * T(0.01506)) * sig)[:, None]
    X = (np.asarray(vertices[M:], T) @ bq.T).astype(T)
    return _sinc(X)


def _peaks_to_outputs(odf, vertices, nbr, T):
    """Shared tail of gqi_rec / dsi_rec voxel loop (src/gqi.jl:147-159)."""
    nv = odf.shape[0]
    odfmin = odf.min(axis=1)
    idx, nvalid = find_peaks_batch(odf, nbr)
    peak = np.zeros((NPEAK, nv, 3), T); qa = np.zeros((NPEAK, nv), T)
    for k in range(NPEAK):
        ok = idx[:, k] >= 0
        peak[k, ok] = vertices[idx[ok, k]]                       # first-half vertex rows
        qa[k, ok] = odf[ok, idx[ok, k]] - odfmin[ok]
    return idx, nvalid, peak, qa


def _seq_mean(odf, T):
    """mean(odf.vol, dims=4): sequential accumulation over the frame axis in T, / M."""
    acc = np.zeros(odf.shape[0], T)
    for j in range(odf.shape[1]):
        acc = acc + odf[:, j]
    return acc / T(odf.shape[1])


def gqi_rec(dwi, mask, bval, bvec, vertices, faces, sigma=1.25, dtype=np.float32):
    """gqi_rec(dwi, mask, odf_dirs, sigma) -> GQI   (src/gqi.jl:109-171).

    Returns dict: odf [nx,ny,nz,M], peak [3][nx,ny,nz,3], qa [3][nx,ny,nz],
    plus test-only extras: peak_idx [nx,ny,nz,3] (0-based, -1 = none), nvalid,
    computed (bool: voxel passed mask and max(s)>0), odfmax."""
    T = dtype
    shp = dwi.shape[:3]
    nvox = int(np.prod(shp))
    M = vertices.shape[0] // 2
    A = gqi_matrix(bval, bvec, vertices, sigma, T)
    nbr = neighbour_table(fold_faces(faces, M), M)
    m = (np.asarray(mask).reshape(shp) != 0).reshape(-1, order="F")
    S = np.asarray(dwi).reshape((-1, dwi.shape[3]), order="F")[m].astype(T)
    S = np.maximum(S, 0)                                         # :140
    comp = S.max(axis=1) != 0                                    # :142
    S = S[comp]
    odf = (S @ A.T).astype(T)                                    # :144 (sgemv per voxel)
    idx, nvalid, peak, qa = _peaks_to_outputs(odf, np.asarray(vertices, T), nbr, T)
    sel = np.nonzero(m)[0][comp]
    ODF = np.zeros((nvox, M), T); ODF[sel] = odf
    mean = np.zeros(nvox, T); mean[sel] = _seq_mean(odf, T)
    odfmax = mean.max() if nvox else T(0)                        # :164 (over ALL voxels incl. zeros)
    out = {"odf": ODF.reshape(shp + (M,), order="F"), "peak": [], "qa": [], "odfmax": odfmax}
    with np.errstate(all="ignore"):
        for k in range(NPEAK):
            P = np.zeros((nvox, 3), T); P[sel] = peak[k]
            Q = np.zeros(nvox, T); Q[sel] = qa[k]
            Q = Q / odfmax                                       # :166-168
            out["peak"].append(P.reshape(shp + (3,), order="F"))
            out["qa"].append(Q.reshape(shp, order="F"))
    I = -np.ones((nvox, NPEAK), np.int32); I[sel] = idx
    NV = np.zeros(nvox, np.int32); NV[sel] = nvalid
    C = np.zeros(nvox, bool); C[sel] = True
    out["peak_idx"] = I.reshape(shp + (NPEAK,), order="F")
    out["nvalid"] = NV.reshape(shp, order="F"); out["computed"] = C.reshape(shp, order="F")
    return out


# --------------------------------------------------------------------------
# DSI   (src/dsi.jl)
# --------------------------------------------------------------------------
def dsi_work(bval, bvec, vertices, hann_width=32, dtype=np.float32):
    """DSIwork  (src/dsi.jl:59-143).  Returns dict with nfft, iq (int [N,3]),
    iq_sub (1-based grid subscripts), H [nfft^3] window, qr2 [nrad], dqr,
    coords [3, nrad, M] (1-based continuous subscripts)."""
    T = dtype
    bval = np.asarray(bval, np.float32); bvec = np.asarray(bvec, np.float32)
    q = bvec * np.sqrt(bval)[:, None]                            # :62
    bmin = bval.min()
    dq = np.sqrt(bval[bval > bmin].min())                        # :66
    iq = np.round(q / dq).astype(np.int32)                       # :67 (round half even, as Julia)
    nfft = int(iq.max() - iq.min() + 1)                          # :70
    nfft = 2 ** int(np.ceil(np.log2(nfft)))                      # :71
    shift = nfft // 2 + 1                                        # :73
    iq_sub = iq + shift                                          # 1-based subscripts
    H = np.zeros((nfft, nfft, nfft), T)
    if hann_width == 0:
        H[:] = 1                                                 # :81
    else:
        # computed in Float64 then rounded on assignment (:84)
        h = (1 + np.cos(np.sqrt((iq.astype(np.float64) ** 2).sum(axis=1)) * (2 * np.pi / hann_width))) * 0.5
        H[iq_sub[:, 0] - 1, iq_sub[:, 1] - 1, iq_sub[:, 2] - 1] = h.astype(T)
    M = vertices.shape[0] // 2
    r32 = np.asarray([0.3 + 0.03 * i for i in range(21)], np.float64)   # 0.3:0.03:0.9 (21 values)
    if T == np.float32:
        qr = (np.float32(nfft / 2 - 1) * r32.astype(np.float32)).astype(np.float32)   # :104
    else:
        qr = T(nfft / 2 - 1) * r32.astype(np.float32).astype(T)
    dqr = T(qr[1] - qr[0])                                       # :105 (fp32 difference)
    V2 = np.asarray(vertices[M:], T)
    coords = (V2.T[:, None, :] * qr[None, :, None] + T(shift)).astype(T)   # [3, nrad, M]  :106-109
    return dict(nfft=nfft, iq=iq, iq_sub=iq_sub, H=H, qr2=(qr ** 2).astype(T), dqr=dqr,
                coords=coords, shift=shift, M=M)


def _trilinear(p, coords, T):
    """Interpolations.jl BSpline(Linear()) evaluation of p [nv, n,n,n] at 1-based
    continuous subscripts coords [3, K] -> [nv, K]; x-innermost nesting."""
    x, y, z = coords[0], coords[1], coords[2]
    ix = np.floor(x).astype(np.int64); iy = np.floor(y).astype(np.int64); iz = np.floor(z).astype(np.int64)
    fx = (x - ix.astype(T)).astype(T); fy = (y - iy.astype(T)).astype(T); fz = (z - iz.astype(T)).astype(T)
    ix -= 1; iy -= 1; iz -= 1   # to 0-based
    one = T(1)

    def g(a, b, c):
        return p[:, a, b, c]
    c00 = g(ix, iy, iz) * (one - fx) + g(ix + 1, iy, iz) * fx
    c10 = g(ix, iy + 1, iz) * (one - fx) + g(ix + 1, iy + 1, iz) * fx
    c01 = g(ix, iy, iz + 1) * (one - fx) + g(ix + 1, iy, iz + 1) * fx
    c11 = g(ix, iy + 1, iz + 1) * (one - fx) + g(ix + 1, iy + 1, iz + 1) * fx
    c0 = c00 * (one - fy) + c10 * fy
    c1 = c01 * (one - fy) + c11 * fy
    return (c0 * (one - fz) + c1 * fz).astype(T)


def dsi_rec(dwi, mask, bval, bvec, vertices, faces, hann_width=32, dtype=np.float32, chunk=2048):
    """dsi_rec(dwi, mask, odf_dirs, hann_width) -> DSI   (src/dsi.jl:171-270), FFT form.

    Returns dict: pdf [nx,ny,nz,N], odf [nx,ny,nz,M], peak[3], qa[3] + extras as gqi_rec."""
    T = dtype
    CT = np.complex64 if T == np.float32 else np.complex128
    shp = dwi.shape[:3]
    nvox = int(np.prod(shp)); N = dwi.shape[3]
    W = dsi_work(bval, bvec, vertices, hann_width, T)
    nfft, M = W["nfft"], W["M"]
    nbr = neighbour_table(fold_faces(faces, M), M)
    sub = W["iq_sub"] - 1
    m = (np.asarray(mask).reshape(shp) != 0).reshape(-1, order="F")
    S = np.asarray(dwi).reshape((-1, N), order="F")[m].astype(T)
    nv = S.shape[0]
    X = np.zeros((nv, nfft, nfft, nfft), T)
    for j in range(N):                                           # :205, duplicates: last write wins
        X[:, sub[j, 0], sub[j, 1], sub[j, 2]] = S[:, j]
    comp = X.reshape(nv, -1).max(axis=1) != 0                    # :207
    X = np.maximum(X[comp], 0) * W["H"][None]                    # :209-212
    nc = X.shape[0]
    odf = np.zeros((nc, M), T); pdf = np.zeros((nc, N), T)
    coords = W["coords"].reshape(3, -1)                          # [3, nrad*M]
    nrad = W["qr2"].shape[0]
    for c0 in range(0, nc, chunk):
        xs = X[c0:c0 + chunk]
        x = np.fft.fftshift(np.fft.fftn(np.fft.fftshift(xs.astype(CT), axes=(1, 2, 3)), axes=(1, 2, 3)),
                            axes=(1, 2, 3))                      # :218-220
        p = np.real(x).astype(T)                                 # :224
        with np.errstate(all="ignore"):
            p = (p / p.reshape(p.shape[0], -1).sum(axis=1, dtype=T)[:, None, None, None]).astype(T)  # :225
        pdf[c0:c0 + chunk] = p[:, sub[:, 0], sub[:, 1], sub[:, 2]]   # :227
        val = _trilinear(p, coords, T).reshape(p.shape[0], nrad, M)   # :236-238
        acc = np.zeros((p.shape[0], M), T)
        for r in range(nrad):
            acc = acc + val[:, r, :] * W["qr2"][r]
        odf[c0:c0 + chunk] = acc * W["dqr"]                      # :241
    idx, nvalid, peak, qa = _peaks_to_outputs(odf, np.asarray(vertices, T), nbr, T)
    sel = np.nonzero(m)[0][comp]
    ODF = np.zeros((nvox, M), T); ODF[sel] = odf
    PDF = np.zeros((nvox, N), T); PDF[sel] = pdf
    mean = np.zeros(nvox, T)
    with np.errstate(all="ignore"):
        mean[sel] = _seq_mean(odf, T)
        odfmax = np.max(mean) if nvox else T(0)
    out = {"pdf": PDF.reshape(shp + (N,), order="F"), "odf": ODF.reshape(shp + (M,), order="F"),
           "peak": [], "qa": [], "odfmax": odfmax}
    with np.errstate(all="ignore"):
        for k in range(NPEAK):
            P = np.zeros((nvox, 3), T); P[sel] = peak[k]
            Q = np.zeros(nvox, T); Q[sel] = qa[k]
            out["peak"].append(P.reshape(shp + (3,), order="F"))
            out["qa"].append((Q / odfmax).reshape(shp, order="F"))
    I = -np.ones((nvox, NPEAK), np.int32); I[sel] = idx
    C = np.zeros(nvox, bool); C[sel] = True
    NV = np.zeros(nvox, np.int32); NV[sel] = nvalid
    out["peak_idx"] = I.reshape(shp + (NPEAK,), order="F")
    out["nvalid"] = NV.reshape(shp, order="F"); out["computed"] = C.reshape(shp, order="F")
    return out


def dsi_matrices(bval, bvec, vertices, hann_width=32):
    """Matrix form of the per-voxel DSI pipeline (float64):
       odf = (Mo @ s+) / den,  pdf = (Mp @ s+) / den,  den = nfft^3 * H_c * s+_c
    where c is the LAST volume mapped to the grid centre.  Columns of volumes that are
    overwritten by a later duplicate q-point are zero.  Returns (Mo [M,N], Mp [N,N],
    cvol, dscale = nfft^3 * H_c).  Derived from src/dsi.jl:205-242; checked against
    `dsi_rec` (FFT form) in tests/test_oracle.py."""
    W = dsi_work(bval, bvec, vertices, hann_width, np.float64)
    nfft, M = W["nfft"], W["M"]
    iq = W["iq"].astype(np.int64); N = iq.shape[0]
    sub = W["iq_sub"] - 1
    lin = (sub[:, 0] * nfft + sub[:, 1]) * nfft + sub[:, 2]
    live = np.ones(N, bool)
    last = {}
    for j in range(N):
        if lin[j] in last:
            live[last[lin[j]]] = False
        last[lin[j]] = j
    Hj = W["H"][sub[:, 0], sub[:, 1], sub[:, 2]] * live
    # P[g] = sum_j H_j s_j cos(2 pi (g - centre).iq_j / nfft)  for grid point g (0-based)
    def prop(points):     # points [..., 3] 0-based grid subscripts -> [..., N]
        rel = points.astype(np.float64) - (W["shift"] - 1)
        return np.cos(2 * np.pi * (rel @ iq.T.astype(np.float64)) / nfft) * Hj
    Mp = prop(sub)
    # (rows of Mp are indexed by grid cell, so an overwritten volume's pdf row equals the
    #  row of the later duplicate, exactly as `p[iq_ind]` at src/dsi.jl:227 repeats the value)
    coords = W["coords"]     # [3, nrad, M], 1-based
    nrad = coords.shape[1]
    Mo = np.zeros((M, N))
    c0 = np.floor(coords).astype(np.int64)
    fr = coords - c0
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                w = (fr[0] if dx else 1 - fr[0]) * (fr[1] if dy else 1 - fr[1]) * (fr[2] if dz else 1 - fr[2])
                pts = np.stack([c0[0] + dx - 1, c0[1] + dy - 1, c0[2] + dz - 1], axis=-1)   # [nrad, M, 3]
                Mo += np.einsum("rm,rmn->mn", w * W["qr2"][:, None], prop(pts))
    Mo *= W["dqr"]
    centre = (W["shift"] - 1) * (nfft * nfft + nfft + 1)
    cvol = last.get(centre, -1)
    dscale = float(nfft ** 3 * (W["H"].reshape(-1)[centre])) if cvol >= 0 else 0.0
    return Mo, Mp, cvol, dscale
