"""Seeded synthetic DWI phantoms for the five BASELINE.json configs.

Host-side (numpy) generators used by the tests and by small bench legs; the
full-size bench volumes are generated on the device in bench.py with the same
signal model.  Nothing here is on the reconstruction hot path.
"""
from __future__ import annotations

import numpy as np


def fibonacci_sphere(n: int, rot: float = 0.0) -> np.ndarray:
    """n roughly uniform unit vectors (rows), rotated about z by `rot` radians."""
    i = np.arange(n) + 0.5
    z = 1 - 2 * i / n
    th = np.pi * (1 + 5 ** 0.5) * i + rot
    r = np.sqrt(1 - z * z)
    v = np.stack([r * np.cos(th), r * np.sin(th), z], axis=1)
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def shells_table(nb0: int, shells: list[tuple[float, int]]):
    """(bval float32 [N], bvec float32 [N,3]); b0 rows first with zero bvec, as
    mri_read leaves them (reference: src/mri.jl:711-712)."""
    bval = [0.0] * nb0
    bvec = [np.zeros((nb0, 3))]
    for k, (b, n) in enumerate(shells):
        bval += [b] * n
        bvec.append(fibonacci_sphere(n, rot=0.7 * k))
    return np.asarray(bval, np.float32), np.concatenate(bvec).astype(np.float32)


def dsi_grid_table(rmax2: int = 25, bunit: float = 400.0):
    """515-point Cartesian q-space grid: all integer (i,j,k) with i²+j²+k² <= 25.
    bval = 400·|q|² (b_max = 10 000), bvec = q/|q| (origin = b0, zero bvec)."""
    r = int(np.floor(np.sqrt(rmax2)))
    g = np.arange(-r, r + 1)
    q = np.array([(i, j, k) for i in g for j in g for k in g if i * i + j * j + k * k <= rmax2], np.float64)
    # b0 first (a reader's usual order), the rest by radius
    order = np.argsort((q ** 2).sum(axis=1), kind="stable")
    q = q[order]
    n2 = (q ** 2).sum(axis=1)
    bval = (bunit * n2).astype(np.float32)
    with np.errstate(all="ignore"):
        bvec = np.where(n2[:, None] > 0, q / np.sqrt(n2)[:, None], 0.0).astype(np.float32)
    return bval, bvec


def _random_dirs(rng, n):
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def ellipsoid_mask(shape, fill=0.55):
    """uint8 ellipsoid mask with approximately `fill` volume fraction."""
    nx, ny, nz = shape
    x, y, z = np.meshgrid(np.linspace(-1, 1, nx), np.linspace(-1, 1, ny), np.linspace(-1, 1, nz), indexing="ij")
    r = (fill * 8 / (4 / 3 * np.pi)) ** (1 / 3)
    return np.asfortranarray(((x * x + y * y + z * z) <= r * r).astype(np.uint8))


def rician(rng, s, sigma):
    return np.sqrt((s + rng.normal(scale=1.0, size=s.shape) * sigma) ** 2 +
                   (rng.normal(scale=1.0, size=s.shape) * sigma) ** 2)


def dti_phantom(shape=(64, 64, 40), nb0=1, ndir=30, b=1000.0, seed=1, snr=30.0, inject=True):
    """cfg1: single-tensor field, Rician noise, a few non-positive samples injected.
    Returns dict(dwi [nx,ny,nz,N] float32 F-order, mask uint8, bval, bvec, truth...)."""
    rng = np.random.default_rng(seed)
    bval, bvec = shells_table(nb0, [(b, ndir)])
    nv = int(np.prod(shape))
    l1 = rng.uniform(0.8e-3, 2.0e-3, nv)
    l23 = np.sort(rng.uniform(0.1e-3, 0.8e-3, (nv, 2)), axis=1)[:, ::-1]
    e1 = _random_dirs(rng, nv)
    t = _random_dirs(rng, nv)
    e2 = np.cross(e1, t); e2 /= np.linalg.norm(e2, axis=1, keepdims=True)
    e3 = np.cross(e1, e2)
    S0 = rng.uniform(500, 1500, nv)
    g = bvec.astype(np.float64)
    adc = (l1[:, None] * (e1 @ g.T) ** 2 + l23[:, :1] * (e2 @ g.T) ** 2 + l23[:, 1:] * (e3 @ g.T) ** 2)
    S = S0[:, None] * np.exp(-bval[None, :].astype(np.float64) * adc)
    if snr > 0:
        S = rician(rng, S, (S0 / snr)[:, None])
    if inject:
        k = rng.choice(nv, max(1, nv // 200), replace=False)          # 0.5 %: 1-3 non-positive samples
        for i in k:
            j = rng.choice(np.arange(nb0, bval.shape[0]), rng.integers(1, 4), replace=False)
            S[i, j] = -np.abs(S[i, j]) * rng.integers(0, 2)
        k = rng.choice(nv, max(1, nv // 1000), replace=False)         # 0.1 %: b0 <= 0
        S[k, :nb0] = 0
        k = rng.choice(nv, max(1, nv // 1000), replace=False)         # 0.1 %: all zero
        S[k, :] = 0
    dwi = np.asfortranarray(S.astype(np.float32).reshape(shape + (bval.shape[0],), order="F"))
    return dict(dwi=dwi, mask=ellipsoid_mask(shape, 0.55), bval=bval, bvec=bvec,
                l1=l1, l2=l23[:, 0], l3=l23[:, 1], e1=e1, S0=S0)


def multifibre_signal(rng, nv, bval, bvec, snr=30.0, lpar=1.7e-3, lperp=2.0e-4, diso=3.0e-3, fiso=0.1):
    """Two-fibre multi-tensor + isotropic compartment, S0 in U(500,1500), Rician noise."""
    f1 = rng.uniform(0.3, 0.7, nv) * (1 - fiso)
    f2 = (1 - fiso) - f1
    e1 = _random_dirs(rng, nv); e2 = _random_dirs(rng, nv)
    S0 = rng.uniform(500, 1500, nv)
    g = bvec.astype(np.float64); b = bval.astype(np.float64)[None, :]
    S = (f1[:, None] * np.exp(-b * (lperp + (lpar - lperp) * (e1 @ g.T) ** 2)) +
         f2[:, None] * np.exp(-b * (lperp + (lpar - lperp) * (e2 @ g.T) ** 2)) +
         fiso * np.exp(-b * diso)) * S0[:, None]
    if snr > 0:
        S = rician(rng, S, (S0 / snr)[:, None])
    return S, e1, e2, f1, f2


def gqi_phantom(shape=(24, 20, 12), nb0=18, shells=((1000.0, 90), (2000.0, 90), (3000.0, 90)),
                seed=2, snr=30.0, mask_fill=None, neg_frac=1e-3):
    """cfg2-shaped (scaled down by default): 18 b0 + 90x(1000,2000,3000) = 288 volumes."""
    rng = np.random.default_rng(seed)
    bval, bvec = shells_table(nb0, list(shells))
    nv = int(np.prod(shape))
    S, e1, e2, f1, f2 = multifibre_signal(rng, nv, bval, bvec, snr)
    if neg_frac > 0:
        neg = rng.random(S.shape) < neg_frac
        S[neg] = -S[neg] * 0.1
        k = rng.choice(nv, max(1, nv // 500), replace=False)
        S[k] = -np.abs(S[k]) * rng.integers(0, 2, (k.shape[0], 1))    # all non-positive voxels (skipped)
    dwi = np.asfortranarray(S.astype(np.float32).reshape(shape + (bval.shape[0],), order="F"))
    mask = np.asfortranarray(np.ones(shape, np.uint8)) if mask_fill is None else ellipsoid_mask(shape, mask_fill)
    return dict(dwi=dwi, mask=mask, bval=bval, bvec=bvec, e1=e1, e2=e2, f1=f1, f2=f2)


def dsi_phantom(shape=(12, 10, 6), seed=3, snr=30.0, mask_fill=None):
    """cfg3-shaped (scaled down by default): 515-point grid, same fibre model."""
    rng = np.random.default_rng(seed)
    bval, bvec = dsi_grid_table()
    nv = int(np.prod(shape))
    S, e1, e2, f1, f2 = multifibre_signal(rng, nv, bval, bvec, snr, lpar=1.2e-3, lperp=1.5e-4, diso=2.0e-3)
    k = rng.choice(nv, max(1, nv // 300), replace=False)
    S[k] = 0
    dwi = np.asfortranarray(S.astype(np.float32).reshape(shape + (bval.shape[0],), order="F"))
    mask = np.asfortranarray(np.ones(shape, np.uint8)) if mask_fill is None else ellipsoid_mask(shape, mask_fill)
    return dict(dwi=dwi, mask=mask, bval=bval, bvec=bvec, e1=e1, e2=e2)
