"""CPU ORACLE (test infrastructure, NOT the product) for RUMBA-SD, SURVEY.md section 8(f) rank 2.

numpy restatement of the reference algorithm, function by function:
    besseli_ratio        /root/reference/src/rusd.jl:167-175   (Perron continued fraction)
    sd_grad! / sd_div!   :181-205
    rumba_tv!            :214-234
    rumba_sd_initialize! :240-255
    rumba_sd_iterate!    :266-340
    rumba_peaks!         :348-373
    rumba_rec            :419-636  (signal normalisation :448-466, neighbourhoods :478-492, kernel :497-520,
                                    start value :527-529, normalisation / GFA / peaks :560-633)
    ang2rot, cart2sph    /root/reference/src/util.jl:48-56, :85-100;  tensor_model src/rusd.jl:141-153

Only tests/, __graft_entry__.smoke() and bench legs labelled cpu_baseline may import this module.
PARITY UNPINNED: the reference ships no test vectors and Julia is not installed (see fibers_oracle.py); third-party
arithmetic restated from its published behaviour: LinearAlgebra.mul! (sgemm), Statistics.mean / std (corrected,
N-1), Base.sum(dims=1), sortperm!(rev=true) (stable).  `dtype=np.float32` follows the reference's Float32 arithmetic
operation by operation (matrix products in float32 via numpy / BLAS, whose summation order differs from OpenBLAS
inside Julia); `dtype=np.float64` is the truth the tolerances are applied against.
"""
from __future__ import annotations

import numpy as np


def besseli_ratio(nu: int, z):
    """I_nu(z) / I_{nu-1}(z) by Perron's continued fraction, exactly the reference's expression (rusd.jl:169-174)."""
    T = z.dtype.type
    n = T(2 * nu)
    return z / ((n + z) - ((n + T(1)) * z / (T(2) * z + (n + T(1)) - ((n + T(3)) * z / ((n + T(2)) + T(2) * z - ((n + T(5)) * z / ((n + T(3)) + T(2) * z)))))))


def ang2rot(phi, theta):
    Rz = np.array([[np.cos(phi), -np.sin(phi), 0], [np.sin(phi), np.cos(phi), 0], [0, 0, 1]])
    Ry = np.array([[np.cos(theta), 0, np.sin(theta)], [0, 1, 0], [-np.sin(theta), 0, np.cos(theta)]])
    return Rz @ Ry


def tensor_model(phi, theta, lam, b, g, s0=1.0):
    R = ang2rot(phi, theta)
    D = R @ np.diag(lam) @ R.T
    return s0 * np.exp(-b * np.einsum("ij,jk,ik->i", g, D, g))


def rumba_kernel(bval, bvec, vertices, lam_para=1.7e-3, lam_perp=0.2e-3, lam_csf=3.0e-3, lam_gm=0.8e-4, dtype=np.float32):
    """(Kernel [ndir, nvert+2], ib0, b, g): rusd.jl:444-520.  Computed in float64, rounded to `dtype`."""
    bval = np.asarray(bval, np.float64); bvec = np.asarray(bvec, np.float64)
    ib0 = bval == bval.min()
    gd = bvec[~ib0]
    with np.errstate(all="ignore"):
        g = np.vstack([np.zeros((1, 3)), gd / np.sqrt((gd ** 2).sum(axis=1, keepdims=True))])
    b = np.concatenate([[0.0], bval[~ib0]])
    nvert = vertices.shape[0] // 2
    v2 = np.asarray(vertices[nvert:], np.float64)
    hyp = np.hypot(v2[:, 0], v2[:, 1])
    theta = -np.arctan2(v2[:, 2], hyp)
    phi = np.arctan2(v2[:, 1], v2[:, 0])
    K = np.zeros((b.shape[0], nvert + 2))
    for i in range(nvert):
        K[:, i] = tensor_model(phi[i], theta[i], [lam_para, lam_perp, lam_perp], b, g)
    K[:, nvert] = tensor_model(0.0, 0.0, [lam_csf] * 3, b, g)
    K[:, nvert + 1] = tensor_model(0.0, 0.0, [lam_gm] * 3, b, g)
    return K.astype(dtype), ib0, b, g


def angular_neighbours(vertices, ang_neig):
    """idx_neig (rusd.jl:478-492) as a boolean [nvert, nvert] matrix (antipodally folded angle < ang_neig, no self)."""
    nvert = vertices.shape[0] // 2
    hv = np.asarray(vertices[:nvert], np.float32)
    c = np.clip(hv @ hv.T, -1, 1)
    ang = np.degrees(np.arccos(c))
    ang = np.minimum(ang, 180 - ang)
    nb = ang < ang_neig
    np.fill_diagonal(nb, False)
    return nb


def default_ang_neig(nvert2):
    return {724: 12.5, 642: 12.5, 362: 16.0}[int(nvert2)]


def signal_matrix(dwi, mask, ib0, dtype=np.float32):
    """signal_mat [ndir, nmask] (rusd.jl:448-466) and ind_mask (0-based linear indices, column-major)."""
    T = dtype
    nx, ny, nz, nvol = dwi.shape
    flat = np.asarray(dwi, T).reshape(-1, nvol, order="F")
    ind = np.nonzero(np.asarray(mask).reshape(-1, order="F") > 0)[0]
    S = np.maximum(flat[ind], T(0))
    s0 = S[:, ib0].mean(axis=1, dtype=T)
    sig = np.empty((int((~ib0).sum()) + 1, ind.shape[0]), T)
    with np.errstate(all="ignore"):
        sig[1:] = (S[:, ~ib0] / s0[:, None]).T
    sig[0] = s0
    sig[np.isnan(sig)] = 0
    sig[0] = (sig[0] > 0).astype(T)
    sig[sig > 1] = 1
    return sig, ind


def _tv_term(V, lam, eps):
    """rumba_tv! on zero-embedded component volumes V [nx,ny,nz,ncomp] (all components at once); lam [nx,ny,nz,1]."""
    Gx = np.concatenate([V[1:], V[-1:]], axis=0) - V
    Gy = np.concatenate([V[:, 1:], V[:, -1:]], axis=1) - V
    Gz = np.concatenate([V[:, :, 1:], V[:, :, -1:]], axis=2) - V
    n = np.sqrt(Gx ** 2 + Gy ** 2 + Gz ** 2 + eps)
    Gx = Gx / n; Gy = Gy / n; Gz = Gz / n
    D = np.zeros_like(V)
    D[1:-1] = Gx[1:-1] - Gx[:-2]; D[0] = Gx[0]; D[-1] = -Gx[-2]
    D[:, 1:-1] += Gy[:, 1:-1] - Gy[:, :-2]; D[:, 0] += Gy[:, 0]; D[:, -1] += -Gy[:, -2]
    D[:, :, 1:-1] += Gz[:, :, 1:-1] - Gz[:, :, :-2]; D[:, :, 0] += Gz[:, :, 0]; D[:, :, -1] += -Gz[:, :, -2]
    return 1 / (np.abs(1 - lam * D) + eps)


def rumba_rec(dwi, mask, bval, bvec, vertices, niter=600, lam_para=1.7e-3, lam_perp=0.2e-3, lam_csf=3.0e-3, lam_gm=0.8e-4,
              ncoils=1, coil_combine="SMF-SENSE", ipat_factor=1, use_tv=True, ang_neig=None, dtype=np.float32, return_state=False):
    T = dtype
    # eps(T) enters the ALGORITHM (TV normalisation sqrt(|G|^2 + eps), 1 / (|1 - lambda div| + eps), ...): the reference is run on
    # Float32 volumes (dwi::MRI{Array{Float32,4}} from mri_read), so the float64 "truth" keeps Float32's eps and only the
    # rounding of the arithmetic changes.
    eps = T(np.finfo(np.float32).eps)
    n_order = ncoils if coil_combine == "SoS-GRAPPA" else 1
    if coil_combine not in ("SoS-GRAPPA", "SMF-SENSE"):
        raise ValueError("Unknown coil combine mode " + coil_combine)
    if ipat_factor < 1:
        raise ValueError("iPAT factor must be a positive integer")
    nx, ny, nz, nvol = dwi.shape
    shape = (nx, ny, nz)
    K, ib0, b, g = rumba_kernel(bval, bvec, vertices, lam_para, lam_perp, lam_csf, lam_gm, T)
    sig, ind = signal_matrix(dwi, mask, ib0, T)
    ndir, ncomp = K.shape
    nvert = ncomp - 2
    nmask = ind.shape[0]
    ang_neig = default_ang_neig(vertices.shape[0]) if ang_neig is None else ang_neig
    nb = angular_neighbours(vertices, ang_neig)
    f0 = np.ones(ncomp, T); f0 = f0 / T(2 * nvert + 2); f0 = f0 / f0.sum(dtype=T)
    fodf = np.tile(f0[:, None], (1, nmask)).astype(T)
    dodf = np.tile((K @ f0)[:, None], (1, nmask)).astype(T)
    lam0 = T(T(1 / 15) ** 2)
    lam = np.full(shape, lam0, T)
    s2 = np.full((1, nmask), lam0, T)
    dsig = (sig * dodf) / s2
    tv = np.ones((ncomp, nmask), T)
    snr = np.zeros((1, nmask), T)
    for it in range(niter):
        Ir = besseli_ratio(n_order, dsig)
        rl = K.T @ (sig * Ir)
        rl2 = K.T @ dodf
        rl = rl / (rl2 + eps)
        if use_tv:
            V = np.zeros((nx * ny * nz, ncomp), T); V[ind] = fodf.T
            tv = _tv_term(V.reshape(shape + (ncomp,), order="F"), lam[..., None], eps).reshape(-1, ncomp, order="F")[ind].T
        fodf = np.maximum(fodf * rl * tv, T(0))
        dodf = (K @ fodf).astype(T)
        dsig = (sig * dodf) / s2
        Ir = (sig ** 2 + dodf ** 2) / T(2) - (s2 * dsig) * Ir
        s2 = (Ir.sum(axis=0, dtype=T, keepdims=True) / T(n_order * ndir)).astype(T)
        s2 = np.clip(s2, T((1 / 80) ** 2), T((1 / 8) ** 2))
        snr = 1 / np.sqrt(s2)
        if use_tv:
            if ipat_factor == 1:
                lam = np.full(shape, max(s2.mean(dtype=T), T((1 / 30) ** 2)), T)
            else:
                lam = np.zeros(nx * ny * nz, T); lam[ind] = s2[0]; lam = lam.reshape(shape, order="F")
    snr_mean = T(snr.mean(dtype=T)) if niter > 0 else T(0)
    snr_std = T(np.sqrt(((snr - snr_mean) ** 2).sum(dtype=T) / T(max(nmask - 1, 1)))) if niter > 0 else T(0)
    state = dict(fodf_mat=fodf.copy(), dodf=dodf.copy(), s2=s2.copy())
    fodf = fodf / (fodf.sum(axis=0, dtype=T, keepdims=True) + eps)
    nvx = nx * ny * nz
    out_fodf = np.zeros((nvx, nvert), T); fcsf = np.zeros(nvx, T); fgm = np.zeros(nvx, T); var = np.zeros(nvx, T)
    out_fodf[ind] = fodf[:nvert].T
    fcsf[ind] = fodf[nvert]; fgm[ind] = fodf[nvert + 1]
    f_iso = fgm + fcsf
    out_fodf = out_fodf + f_iso[:, None]
    with np.errstate(all="ignore"):
        out_fodf = out_fodf / out_fodf.sum(axis=1, dtype=T, keepdims=True)
        out_fodf[np.isnan(out_fodf)] = 0
        var[ind] = s2[0]
        mean = out_fodf.mean(axis=1, dtype=T)
        std = np.sqrt(((out_fodf - mean[:, None]) ** 2).sum(axis=1, dtype=T) / T(nvert - 1))
        gfa = std / np.sqrt((out_fodf ** 2).mean(axis=1, dtype=T))
        gfa[np.isnan(gfa)] = 0
    npeak, fthresh = 5, T(0.1)
    peaks = np.zeros((npeak, nvx, 3), T); peak_idx = np.full((nvx, npeak), -1, np.int32)
    hv = np.asarray(vertices[:nvert], T)
    mflat = np.asarray(mask).reshape(-1, order="F") != 0
    nbi = [np.nonzero(nb[i])[0] for i in range(nvert)]
    with np.errstate(all="ignore"):
        for v in np.nonzero(mflat)[0]:
            f = out_fodf[v]
            thr_abs = (fthresh / (T(1) - f_iso[v])) * f.max()
            pk = f.copy()
            for i in range(nvert):
                if f[i] < thr_abs or f[i] <= f[nbi[i]].max():
                    pk[i] = 0
            order = np.argsort(-pk, kind="stable")
            n = min(int((pk > 0).sum()), npeak)
            fnorm = (T(1) - f_iso[v]) / f[order[:n]].sum(dtype=T) if n > 0 else T(0)
            for k in range(n):
                peaks[k, v] = hv[order[k]] * (f[order[k]] * fnorm)
                peak_idx[v, k] = order[k]
    res = dict(fodf=out_fodf.reshape((nx, ny, nz, nvert), order="F"), fgm=fgm.reshape(shape, order="F"), fcsf=fcsf.reshape(shape, order="F"),
               peak=[peaks[k].reshape((nx, ny, nz, 3), order="F") for k in range(npeak)], gfa=gfa.reshape(shape, order="F"),
               var=var.reshape(shape, order="F"), snr_mean=float(snr_mean), snr_std=float(snr_std),
               peak_idx=peak_idx.reshape((nx, ny, nz, npeak), order="F"))
    if return_state:
        res["state"] = state
    return res
