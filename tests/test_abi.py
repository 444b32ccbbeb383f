"""CPU tests of the C-ABI boundary and the host-side logic of libfibers_cuda.so (no compute calls:
there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import fibers_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from fibers_jl_b200 import _lib
    return _lib.lib()


def test_library_exports_every_declared_symbol(L):
    from fibers_jl_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "fibers_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fibers_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/fibers_cuda.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes binding and header must list the same entry points"
    assert L.fibers_cuda_version() == 100


def test_no_cpu_fallback_without_device():
    import fibers_jl_b200 as F
    if F.device_count() > 0:
        pytest.skip("a CUDA device is present")
    from fibers_jl_b200 import phantom
    ph = phantom.dti_phantom((4, 4, 2), seed=1)
    with pytest.raises(F.FibersCudaError) as e:
        F.dti_fit(F.MRI(ph["dwi"], ph["bval"], ph["bvec"]), F.MRI(ph["mask"]))
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)
    with pytest.raises(F.FibersCudaError):
        F.gqi_rec(F.MRI(ph["dwi"], ph["bval"], ph["bvec"]), F.MRI(ph["mask"]))
    with pytest.raises(F.FibersCudaError) as e2:                     # the fused DTI + GQI entry point refuses as well
        F.dti_gqi_fit(F.MRI(ph["dwi"], ph["bval"], ph["bvec"]), F.MRI(ph["mask"]))
    assert e2.value.code == 3
    # plan creation (device API) also refuses
    from fibers_jl_b200 import _lib
    plan = C.c_void_p()
    rc = _lib.lib().fibers_dti_plan_create(C.byref(plan), 0, 31, _lib.ptr(ph["bval"]), _lib.ptr(np.asfortranarray(ph["bvec"])))
    assert rc == 3 and not plan.value


def test_reference_error_behaviour():
    import fibers_jl_b200 as F
    m = F.MRI(np.ones((2, 2, 2), np.uint8))
    with pytest.raises(RuntimeError, match="Missing b-value table from input DWI structure"):
        F.adc_fit(F.MRI(np.zeros((2, 2, 2, 4), np.float32)), m)
    with pytest.raises(RuntimeError, match="Missing gradient table from input DWI structure"):
        F.dsi_rec(F.MRI(np.zeros((2, 2, 2, 4), np.float32), bval=np.ones(4, np.float32)), m)
    with pytest.raises(TypeError):      # reference: MethodError for non-Float32 DWI in dti_fit
        F.dti_fit(F.MRI(np.zeros((2, 2, 2, 4), np.int16), np.ones(4, np.float32), np.ones((4, 3), np.float32)), m)
    with pytest.raises(TypeError):      # ... which the fused call inherits
        F.dti_gqi_fit(F.MRI(np.zeros((2, 2, 2, 4), np.int16), np.ones(4, np.float32), np.ones((4, 3), np.float32)), m)
    with pytest.raises(RuntimeError, match="Missing b-value table from input DWI structure"):
        F.dti_gqi_fit(F.MRI(np.zeros((2, 2, 2, 4), np.float32)), m)


def test_mri_container_layout():
    import fibers_jl_b200 as F
    ref = F.MRI(np.zeros((3, 4, 5), np.uint8))
    a = F.MRI.like(ref, 1); b = F.MRI.like(ref, 3)
    assert a.vol.shape == (3, 4, 5) and b.vol.shape == (3, 4, 5, 3)          # nframes == 1 -> 3-D (src/mri.jl:251-255)
    assert b.vol.flags.f_contiguous and b.vol.dtype == np.float32 and not b.vol.any()
    assert F.sphere_642.nvert == 321 and F.sphere_362.nvert == 181 and F.sphere_724.nvert == 362


def _build(L, kind, bval, bvec, v=None, sigma=0.0, hann=0):
    from fibers_jl_b200 import _lib
    n = bval.shape[0]
    M = 0 if v is None else v.shape[0] // 2
    rows = {1: 7, 2: 2, 3: M, 4: M + n}[kind]
    out = np.zeros((rows, n), np.float32)
    cv, ds = C.c_int(-9), C.c_float(0)
    V = None if v is None else np.asfortranarray(v)
    r = L.fibers_host_build_matrix(kind, n, _lib.ptr(bval), _lib.ptr(np.asfortranarray(bvec)), _lib.ptr(V) if V is not None else None,
                                   0 if v is None else v.shape[0], sigma, hann, _lib.ptr(out), out.size, C.byref(cv), C.byref(ds))
    assert r == rows, L.fibers_cuda_last_error()
    return out, cv.value, ds.value


def test_host_matrices_match_oracle(L, sphere642):
    from fibers_jl_b200 import phantom
    v, f = sphere642
    bval, bvec = phantom.shells_table(18, [(1000.0, 90), (2000.0, 90), (3000.0, 90)])
    A, _, _ = _build(L, 3, bval, bvec, v, sigma=1.25)
    assert np.abs(A - O.gqi_matrix(bval, bvec, v, 1.25, np.float64)).max() < 5e-7
    assert np.all(A[:, :18] == 1)
    pA, _, _ = _build(L, 1, bval, bvec)
    _, _, ref = O.dti_design(bval, bvec, np.float64)
    assert np.abs(pA - ref).max() / np.abs(ref).max() < 1e-6
    pB, _, _ = _build(L, 2, bval, bvec)
    _, _, ref = O.adc_design(bval, np.float64)
    assert np.abs(pB - ref).max() / np.abs(ref).max() < 1e-6
    bval, bvec = phantom.dsi_grid_table()
    for hann in (32, 0):
        MM, cvol, dscale = _build(L, 4, bval, bvec, v, hann=hann)
        Mo, Mp, c0, d0 = O.dsi_matrices(bval, bvec, v, hann)
        assert (cvol, dscale) == (c0, d0)
        assert np.abs(MM[:321] - Mo).max() / np.abs(Mo).max() < 2e-6
        assert np.abs(MM[321:] - Mp).max() < 2e-7


@pytest.mark.parametrize("n", [362, 642, 724])
def test_host_neighbour_table(L, n):
    from fibers_jl_b200 import _lib
    v, f = O.load_sphere(n)
    M = n // 2
    nb = np.zeros((M, 8), np.uint16)
    assert L.fibers_host_build_neighbours(_lib.ptr(np.asfortranarray(f)), f.shape[0], M, _lib.ptr(nb)) == 0
    ref = O.neighbour_table(O.fold_faces(f, M), M)
    assert np.array_equal(np.where(nb == 0xFFFF, -1, nb.astype(np.int32)), ref)
    bad = np.asfortranarray(f.copy()); bad[0, 0] = n + 5
    assert L.fibers_host_build_neighbours(_lib.ptr(bad), f.shape[0], M, _lib.ptr(nb)) == 1


def test_slab_partitioner(L):
    from fibers_jl_b200 import _lib, phantom
    mask = phantom.ellipsoid_mask((20, 18, 31), 0.4)
    nxny, nz = 20 * 18, 31
    for ngpu in (1, 2, 4, 8):
        out = np.zeros(2 * ngpu, np.int64)
        assert L.fibers_host_partition_slabs(_lib.ptr(mask), nxny, nz, ngpu, _lib.ptr(out)) == 0
        r = out.reshape(ngpu, 2)
        assert r[0, 0] == 0 and r[-1, 1] == nxny * nz
        assert np.all(r[1:, 0] == r[:-1, 1]), "shards are contiguous and disjoint"
        assert np.all(r % nxny == 0), "shard boundaries are z-slab boundaries"
        assert np.all(r[:, 1] > r[:, 0]), "no empty shard"
        cnt = np.array([mask.reshape(-1, order="F")[a:b].sum() for a, b in r])
        assert cnt.max() <= mask.sum() / ngpu + 2 * mask.reshape(nxny, nz, order="F").sum(axis=0).max()
