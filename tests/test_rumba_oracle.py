"""CPU tests of the RUMBA-SD oracle (known answers) and of the library's host-side RUMBA set-up against it."""
import numpy as np
import pytest

import fibers_oracle as O
import rumba_oracle as R


def test_besseli_ratio_against_scipy():
    """Perron's continued fraction (src/rusd.jl:167-175) against the exponentially scaled Bessel functions."""
    from scipy.special import ive
    z = np.concatenate([np.linspace(1e-3, 5, 200), np.linspace(5, 400, 300)])
    for nu in (1, 4):
        got = R.besseli_ratio(nu, z.astype(np.float64))
        want = ive(nu, z) / ive(nu - 1, z)
        # the reference truncates the fraction after four terms: ~1 % off near z ~ 3, 2e-5 at z = 50, < 1e-6 beyond z = 200
        # (z = signal * dodf / sigma^2 is 64 .. 6400 for a unit signal at the clamped SNR range 8 .. 80)
        assert np.abs(got - want).max() < 1e-2
        assert np.abs(got - want)[z > 50].max() < 2e-5 and np.abs(got - want)[z > 200].max() < 1e-6
    z32 = z.astype(np.float32)
    assert np.abs(R.besseli_ratio(1, z32) - R.besseli_ratio(1, z.astype(np.float64))).max() < 1e-6


def test_single_fibre_known_answer():
    """Noise-free single fibre + isotropic compartment: the strongest peak lies within one mesh spacing of the
    true direction, the volume fractions are a partition of unity."""
    from fibers_jl_b200 import phantom
    v, f = O.load_sphere(724)
    bval, bvec = phantom.shells_table(2, [(2000.0, 60)])
    rng = np.random.default_rng(5)
    shape = (4, 3, 3); nv = int(np.prod(shape))
    e = rng.normal(size=(nv, 3)); e /= np.linalg.norm(e, axis=1, keepdims=True)
    g = bvec.astype(np.float64); b = bval.astype(np.float64)[None, :]
    S = 1000 * (0.8 * np.exp(-b * (0.2e-3 + 1.5e-3 * (e @ g.T) ** 2)) + 0.2 * np.exp(-b * 3.0e-3))
    dwi = np.asfortranarray(S.astype(np.float32).reshape(shape + (bval.shape[0],), order="F"))
    mask = np.ones(shape, np.uint8)
    r = R.rumba_rec(dwi, mask, bval, bvec, v, niter=150, dtype=np.float64)
    pk = r["peak"][0].reshape(-1, 3, order="F")
    pk /= np.linalg.norm(pk, axis=1, keepdims=True)
    assert np.abs((pk * e).sum(axis=1)).min() > np.cos(np.radians(9.0))
    tot = r["fodf"].sum(axis=-1)
    assert np.allclose(tot, 1.0, atol=1e-6)
    assert (r["fcsf"] >= 0).all() and (r["fgm"] >= 0).all() and r["fcsf"].mean() > 0.1          # the isotropic part is found
    assert 8 <= r["snr_mean"] <= 80


@pytest.mark.parametrize("nsphere,ang", [(724, 12.5), (362, 16.0)])
def test_library_host_setup_matches_oracle(nsphere, ang):
    import ctypes as C
    from fibers_jl_b200 import _lib, phantom
    L = _lib.lib()
    v, f = O.load_sphere(nsphere)
    bval, bvec = phantom.shells_table(3, [(1000.0, 20), (3000.0, 30)])
    bval = bval.copy(); bval[:3] = 5.0                                  # minimum b is not zero (HCP-style b = 5)
    nvert = nsphere // 2
    Ko, ib0, b, g = R.rumba_kernel(bval, bvec, v, dtype=np.float64)
    V = np.asfortranarray(v, np.float32); bv = np.asfortranarray(bvec, np.float32)
    K = np.zeros((Ko.shape[0], nvert + 2), np.float32, order="F")
    vr = np.zeros(bval.shape[0], np.int32); nb = np.zeros((nvert, 16), np.uint16)
    ndir = L.fibers_host_build_rumba(bval.shape[0], _lib.ptr(bval), _lib.ptr(bv), _lib.ptr(V), nsphere, ang, 1.7e-3, 0.2e-3, 3.0e-3, 0.8e-4,
                                     _lib.ptr(K), K.size, _lib.ptr(vr), _lib.ptr(nb))
    assert ndir == Ko.shape[0] == int((~ib0).sum()) + 1
    assert np.abs(K - Ko).max() < 2e-7
    assert np.array_equal(vr == 0, ib0) and np.array_equal(vr[~ib0], np.arange(1, ndir))
    want = R.angular_neighbours(v, ang)
    for i in range(nvert):
        got = [int(x) for x in nb[i] if x != 0xFFFF]
        assert got == list(np.nonzero(want[i])[0])
