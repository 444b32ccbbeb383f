"""Structure tensor: oracle known answers on the CPU, GPU parity through the C ABI
(src/structens.jl:13-34 st_eigen, :40-88 st_recon)."""
import numpy as np
import pytest

import structens_oracle as S


def _planes(shape, normal, period=6.0):
    x, y, z = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in shape], indexing="ij")
    n = np.asarray(normal, np.float64); n /= np.linalg.norm(n)
    return np.sin(2 * np.pi * (x * n[0] + y * n[1] + z * n[2]) / period), n


def test_oracle_known_answers():
    g = S.gaussian_taps(1.5)
    assert g.shape[0] == 9 and abs(g.sum() - 1) < 1e-12 and np.allclose(g, g[::-1])
    # a stack of parallel planes: the largest eigenvalue's vector is the plane normal, the other two eigenvalues vanish
    vol, n = _planes((24, 20, 18), (1.0, 2.0, -1.5))
    evec, evals = S.st_recon(vol, 1.0, 2.0)
    c = (slice(6, -6),) * 3
    top = evec[c][..., :, 2]
    assert np.abs(top @ n).min() > 0.999
    assert (evals[c][..., 1] / evals[c][..., 2]).max() < 2e-2
    # mirrored border: a constant volume has zero gradient everywhere, including the border
    _, ev0 = S.st_recon(np.full((6, 5, 4), 3.0), 1.0, 1.0)
    assert np.abs(ev0).max() < 1e-20


@pytest.mark.gpu
def test_st_eigen_and_recon_parity_on_the_gpu():
    import fibers_jl_b200 as F
    import fibers_oracle as O
    assert F.device_count() > 0
    rng = np.random.default_rng(4)
    shape = (21, 17, 13)
    # random symmetric positive semi-definite tensors + special cases (diagonal, zero, repeated eigenvalues)
    G = rng.normal(size=shape + (3, 3))
    A = np.einsum("...ij,...kj->...ik", G, G)
    A[0, 0, 0] = 0; A[1, 0, 0] = np.diag([3.0, 1.0, 2.0]); A[2, 0, 0] = np.eye(3) * 0.5
    comps = [np.asfortranarray(A[..., i, j].astype(np.float32)) for i, j in ((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))]
    evec, evals = F.st_eigen(*comps)
    assert evec.shape == shape + (3, 3) and evals.shape == shape + (3,)
    wv, ww = S.st_eigen(comps)
    scale = np.abs(ww).max(axis=-1, keepdims=True) + 1e-30
    assert np.abs(evals - ww).max() < 1e-30 + 1e-4 * scale.max() and (np.abs(evals - ww) / scale).max() < 1e-4
    assert np.all(np.diff(evals, axis=-1) >= -1e-5 * scale)                                        # ascending
    gap = np.minimum(np.diff(ww, axis=-1)[..., [0, 0, 1]], np.diff(ww, axis=-1)[..., [0, 1, 1]]) / scale   # separation of each value
    dots = np.abs(np.einsum("...ik,...ik->...k", evec.astype(np.float64), wv))
    assert dots[gap > 1e-2].min() >= 0.9999
    # bit-identical to the fp32 closed form the DTI kernel uses (same device function), checked against its numpy restatement
    w32, v32 = O.eig3_sym(*[c.reshape(-1, order="F") for c in comps], dtype=np.float32)
    assert np.abs(evals.reshape(-1, 3, order="F") - w32).max() <= 2e-6 * scale.max()
    # st_recon on a textured volume with two crossing plane families
    v1, n1 = _planes((40, 36, 28), (1.0, 0.3, 0.2)); v2, _ = _planes((40, 36, 28), (-0.2, 1.0, 0.5), 9.0)
    vol = np.asfortranarray((v1 + 0.5 * v2 + 0.05 * rng.normal(size=v1.shape)).astype(np.float32))
    for sigma, rho in ((1.0, 2.0), (0.0, 1.5), (2.5, 0.0)):
        evec, evals = F.st_recon(vol, sigma, rho)
        wv, ww = S.st_recon(vol, sigma, rho)
        sc = np.abs(ww).max()
        # 1e-4 of the voxel's largest eigenvalue where the three values are separated; nearly degenerate pairs (rho = 0 gives
        # rank-1 tensors: two exact zeros) are ill-conditioned for the fp32 closed form itself (up to 5e-4)
        loc = np.abs(ww).max(axis=-1, keepdims=True) + 1e-6 * sc
        sep = (np.diff(ww, axis=-1).min(axis=-1, keepdims=True) / loc) > 1e-2
        e = np.abs(evals - ww) / loc
        assert e[np.broadcast_to(sep, e.shape)].max(initial=0) < 1e-4 and e.max() < 1e-3, (sigma, rho, e.max())
        rel_gap = (ww[..., 2] - ww[..., 1]) / (np.abs(ww[..., 2]) + 1e-30)
        strong = (rel_gap > 0.2) & (ww[..., 2] > 1e-3 * sc)
        d = np.abs(np.einsum("...i,...i->...", evec[..., :, 2].astype(np.float64), wv[..., :, 2]))
        assert strong.mean() > 0.3 and d[strong].min() > 0.9999, (sigma, rho, d[strong].min())
    with pytest.raises(TypeError):
        F.st_recon(vol.astype(np.float64), 1.0, 1.0)


def _golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "structens_small.npz"))


def test_oracle_reproduces_structens_fixture():
    """tests/golden/structens_small.npz (tools/make_golden.py): the fixture the pinning kit hands to the real reference."""
    g = _golden()
    evec, evals = S.st_recon(g["vol"], float(g["sigma"]), float(g["rho"]))
    assert np.array_equal(evals, g["eigval"]) and np.array_equal(evec.astype(np.float32), g["eigvec"])


@pytest.mark.gpu
def test_st_recon_gpu_matches_golden_fixture():
    import fibers_jl_b200 as F
    g = _golden()
    evec, evals = F.st_recon(np.asfortranarray(g["vol"]), float(g["sigma"]), float(g["rho"]))
    ww = g["eigval"]; wv = g["eigvec"].astype(np.float64)
    sc = np.abs(ww).max()
    loc = np.abs(ww).max(axis=-1, keepdims=True) + 1e-6 * sc
    sep = (np.diff(ww, axis=-1).min(axis=-1, keepdims=True) / loc) > 1e-2
    e = np.abs(evals - ww) / loc
    assert e[np.broadcast_to(sep, e.shape)].max(initial=0) < 1e-4 and e.max() < 1e-3
    strong = ((ww[..., 2] - ww[..., 1]) / (np.abs(ww[..., 2]) + 1e-30) > 0.2) & (ww[..., 2] > 1e-3 * sc)
    d = np.abs(np.einsum("...i,...i->...", evec[..., :, 2].astype(np.float64), wv[..., :, 2]))
    assert strong.mean() > 0.3 and d[strong].min() > 0.9999
