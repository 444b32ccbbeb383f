"""CPU ORACLE (test infrastructure, NOT the product) for the structure-tensor path, SURVEY.md section 8(f) rank 3.

    st_eigen  /root/reference/src/structens.jl:13-34   per-voxel eigen(Symmetric(S, :L)): values ascending, vectors in columns
    st_recon  /root/reference/src/structens.jl:40-88   imfilter(gaussian sigma) -> Scharr gradients -> products -> imfilter(gaussian rho)
                                                      -> st_eigen

ImageFiltering.jl is not vendored (Project.toml compat only), so its behaviour is restated from its published definitions:
KernelFactors.gaussian(s) = normalised exp(-x^2 / 2 s^2) on 4 ceil(s) + 1 taps; KernelFactors.scharr = ([-1,0,1] / 2 along the gradient
axis, [3,10,3] / 16 along the others); imfilter correlates; the "reflect" border mirrors about the edge sample without repeating it
(scipy.ndimage mode "mirror").  PARITY UNPINNED (no reference vectors, no Julia).
Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import correlate1d


def gaussian_taps(sigma, dtype=np.float64):
    n = 4 * int(np.ceil(sigma)) + 1
    x = np.arange(n) - n // 2
    w = np.exp(-x * x / (2.0 * sigma * sigma))
    return (w / w.sum()).astype(dtype)


def _sep(vol, fx, fy, fz):
    out = correlate1d(vol, fx.astype(vol.dtype), axis=0, mode="mirror")
    out = correlate1d(out, fy.astype(vol.dtype), axis=1, mode="mirror")
    return correlate1d(out, fz.astype(vol.dtype), axis=2, mode="mirror")


def st_tensor(vol, sigma, rho, dtype=np.float64):
    """The six smoothed products (gxx, gxy, gxz, gyy, gyz, gzz) of src/structens.jl:43-84."""
    img = np.asarray(vol, dtype)
    if sigma > 0:
        g = gaussian_taps(sigma); img = _sep(img, g, g, g)
    der = np.array([-0.5, 0.0, 0.5]); smo = np.array([3.0, 10.0, 3.0]) / 16.0
    gx = _sep(img, der, smo, smo); gy = _sep(img, smo, der, smo); gz = _sep(img, smo, smo, der)
    S = [gx * gx, gx * gy, gx * gz, gy * gy, gy * gz, gz * gz]
    if rho > 0:
        g = gaussian_taps(rho); S = [_sep(s, g, g, g) for s in S]
    return S


def st_eigen(S, dtype=np.float64):
    """(eigvec [nx,ny,nz,3,3], eigval [nx,ny,nz,3]): ascending eigenvalues, eigvec[..., :, k] the k-th vector (sign arbitrary)."""
    sxx, sxy, sxz, syy, syz, szz = [np.asarray(s, np.float64) for s in S]
    A = np.stack([np.stack([sxx, sxy, sxz], -1), np.stack([sxy, syy, syz], -1), np.stack([sxz, syz, szz], -1)], -2)
    w, v = np.linalg.eigh(A)
    return v.astype(dtype), w.astype(dtype)


def st_recon(vol, sigma, rho, dtype=np.float64):
    return st_eigen(st_tensor(vol, sigma, rho, dtype), dtype)
