"""Streamline tractography: known-answer tests of the oracle (the reference's quirks spelled out by hand)
and bit-exact parity of the CUDA path (fibers_stream through the C ABI) against the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import stream_oracle as SO  # noqa: E402

F = np.float32


def straight_field(shape=(20, 5, 5)):
    v = np.zeros(shape + (3,), F, order="F")
    v[..., 0] = 1
    return v


def test_oracle_straight_line_quirks():
    """All vectors along +x in a 20 x 5 x 5 volume, seed voxel (10, 3, 3), no sub-voxel offset, step 0.5, len_max = 20.
    Forward: the positions 10, 10.5, ... are stored while the NEXT position is inside; round(10.5) = 10 and round(20.5) = 20
    (half to even), so 10 ... 20.0 = 21 points are stored and the counter (21 > len_max) stops the pass.  Backward: the seed is
    stored again (counter 22 > len_max) and the pass stops.  Forward points are prepended: 20.0, 19.5, ..., 10.0, then 10.0."""
    v = straight_field()
    m, arr = SO.stream_work([v])
    assert m.all()
    s = SO.new_line([10, 3, 3], np.zeros(3, F), m, arr, len_max=20, cosang_thresh=F(np.cos(np.pi / 4)), step=0.5, smooth=0.2)
    assert s.shape == (3, 22)
    np.testing.assert_array_equal(s[0], np.concatenate([np.arange(20.0, 9.75, -0.5), [10.0]]).astype(F))
    np.testing.assert_array_equal(s[1], np.full(22, 3, F))


def test_oracle_backward_pass_and_bounds():
    """Same field, len_max large: forward runs to the +x face, backward to the -x face; the seed appears twice."""
    v = straight_field()
    m, arr = SO.stream_work([v])
    s = SO.new_line([10, 3, 3], np.zeros(3, F), m, arr, len_max=1000, cosang_thresh=F(0.7071), step=0.5, smooth=0.0)
    x = s[0]
    # forward: 10 ... 20.0 stored (next of 20.5 is 21.0 -> outside); backward: 10, 9.5, ..., 1.0 stored? next of 1.0 is 0.5 -> round = 0 -> outside,
    # so 1.0 is NOT stored; last stored is 1.5 (its next position 1.0 is inside)
    np.testing.assert_array_equal(x, np.concatenate([np.arange(20.0, 9.75, -0.5), np.arange(10.0, 1.25, -0.5)]).astype(F))


def test_oracle_pick_by_angle_sign_and_threshold():
    """Two vector fields: e1 = -x (the propagation flips its sign), e2 = +y.  A line started along e1 keeps to x (|cos| = 1 beats 0).
    A 60-degree kink between two half-volumes stops the line at the 45-degree threshold after the first point past the kink."""
    shape = (12, 12, 3)
    a = np.zeros(shape + (3,), F, order="F"); a[..., 0] = -1
    b = np.zeros(shape + (3,), F, order="F"); b[..., 1] = 1
    m, arr = SO.stream_work([a, b])
    s = SO.new_line([6, 6, 2], np.zeros(3, F), m, arr, len_max=100, cosang_thresh=F(0.70710677), step=0.5, smooth=0.0)
    assert np.all(s[1] == 6) and np.all(s[2] == 2) and s.shape[1] > 10
    k = np.zeros(shape + (3,), F, order="F"); k[:6, :, :, 0] = 1
    k[6:, :, :, 0] = F(np.cos(np.deg2rad(60.0))); k[6:, :, :, 1] = F(np.sin(np.deg2rad(60.0)))
    m, arr = SO.stream_work([k])
    s = SO.new_line([3, 6, 2], np.zeros(3, F), m, arr, len_max=100, cosang_thresh=F(0.70710677), step=0.5, smooth=0.0)
    fwd = s[:, : np.argmax(s[0] == 3.0) + 1][:, ::-1]            # forward part in visiting order
    # 6.0 -> 6.5 stays in voxel 6 (round(6.5) = 6); 6.5 -> 7.0 enters the kinked half: 6.5 is stored, cos 60 < cos 45 ends the pass
    assert fwd[0, -1] == 6.5 and fwd[0, -2] == 6.0
    assert s.shape[1] < 30


def test_oracle_masks():
    v = straight_field((8, 4, 4))
    v[4:, :, :, :] = 0                                          # mask derived from "any component non-zero"
    m, arr = SO.stream_work([v])
    assert m[:4].all() and not m[4:].any()
    fvol = np.ones((8, 4, 4), F, order="F"); fvol[2] = 0.01     # amplitude below f_thresh drops the vector, not the voxel
    m2, arr2 = SO.stream_work([v], f=[fvol], f_thresh=0.03)
    assert m2[2].all() and not arr2[:, 0, 2].any()
    fa = np.full((8, 4, 4), 0.5, F, order="F"); fa[1] = 0.05
    m3, _ = SO.stream_work([v], fa=fa, fa_thresh=0.1)
    assert not m3[1].any() and m3[0].all()
    out = SO.stream([v], [np.zeros(3, F)], len_min=3)
    assert all(s.shape[1] >= 3 for s in out)


def noisy_field(shape, nvec, seed):
    """Smooth random orientation fields with holes (zero vectors) and per-vector amplitudes."""
    g = np.random.default_rng(seed)
    nx, ny, nz = shape
    xs, ys, zs = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    vols, fs = [], []
    for i in range(nvec):
        ph = g.uniform(0, 2 * np.pi, 6)
        th = 0.6 * np.sin(xs / 5.0 + ph[0]) + 0.5 * np.cos(ys / 4.0 + ph[1]) + 0.3 * np.sin(zs / 3.0 + ph[2]) + i * 1.1
        el = 0.5 * np.sin(xs / 6.0 + ph[3]) * np.cos(ys / 7.0 + ph[4]) + 0.2 * np.sin(zs / 2.5 + ph[5])
        v = np.stack([np.cos(th) * np.cos(el), np.sin(th) * np.cos(el), np.sin(el)], axis=-1)
        v += 0.05 * g.standard_normal(v.shape)
        v /= np.linalg.norm(v, axis=-1, keepdims=True)
        hole = g.random(shape) < (0.03 + 0.1 * i)
        v[hole] = 0
        vols.append(np.asfortranarray(v.astype(F)))
        fs.append(np.asfortranarray(g.uniform(0.0, 0.3, shape).astype(F)))
    return vols, fs


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["plain", "f_fa_mask_seed", "nosmooth_single"])
def test_stream_gpu_parity_bit_exact(case):
    import fibers_jl_b200 as Fb
    shape = (26, 22, 12)
    nvec = 1 if case == "nosmooth_single" else 3
    vols, fs = noisy_field(shape, nvec, seed=7)
    g = np.random.default_rng(11)
    sub = Fb.draw_sublist(3, rng=5)
    kw = dict(len_min=3, step_size=0.5, smooth_coeff=0.2, ang_thresh=45)
    okw = dict(len_min=3, step_size=0.5, smooth_coeff=0.2, cosang_thresh=F(np.cos(np.deg2rad(45.0))))
    if case == "f_fa_mask_seed":
        fa = np.asfortranarray(g.uniform(0, 0.6, shape).astype(F))
        mask = np.asfortranarray((g.random(shape) < 0.9).astype(np.uint8))
        seed = np.asfortranarray((g.random(shape) < 0.3).astype(np.uint8))
        kw.update(f=[Fb.MRI(x) for x in fs], f_thresh=0.05, fa=Fb.MRI(fa), fa_thresh=0.1, mask=Fb.MRI(mask), seed=Fb.MRI(seed))
        okw.update(f=fs, f_thresh=0.05, fa=fa, fa_thresh=0.1, mask=mask, seed=seed)
    if case == "nosmooth_single":
        kw.update(smooth_coeff=0.0, step_size=0.75, ang_thresh=30, len_max=9)
        okw.update(smooth_coeff=0.0, step_size=0.75, cosang_thresh=F(np.cos(np.deg2rad(np.float64(F(30))))), len_max=9)
    got = Fb.stream([Fb.MRI(v) for v in vols], sublist=sub, **kw)
    ref = SO.stream(vols, list(sub), **okw)
    assert got.n_count == len(ref) and got.n_count > 200
    assert np.array_equal(got.npts, np.array([s.shape[1] for s in ref], np.int32))
    nbad = sum(0 if np.array_equal(a, b) else 1 for a, b in zip(got.xyz, ref))
    print(f"[parity] stream {case}: {got.n_count} streamlines, {int(got.npts.sum())} points, mismatching lines {nbad}")
    assert nbad == 0


@pytest.mark.gpu
def test_stream_gpu_consumes_gqi_peaks():
    """gqi_rec peaks + qa feed stream() (the reference's downstream consumer, src/stream.jl:76-173 reads peaks / f / fa)."""
    import fibers_jl_b200 as Fb
    from fibers_jl_b200 import phantom
    ph = phantom.gqi_phantom((16, 14, 8), seed=3, mask_fill=0.8)
    r = Fb.gqi_rec(Fb.MRI(ph["dwi"], ph["bval"], ph["bvec"]), Fb.MRI(ph["mask"]))
    sub = Fb.draw_sublist(2, rng=1)
    got = Fb.stream(r.peak, f=r.qa, f_thresh=0.02, mask=Fb.MRI(ph["mask"]), sublist=sub)
    ref = SO.stream([p.vol for p in r.peak], list(sub), f=[q.vol for q in r.qa], f_thresh=0.02, mask=ph["mask"])
    assert got.n_count == len(ref) and got.n_count > 0
    assert all(np.array_equal(a, b) for a, b in zip(got.xyz, ref))


def test_stream_rejects_what_is_not_on_the_gpu_path():
    import fibers_jl_b200 as Fb
    v = straight_field((4, 4, 4))
    with pytest.raises(IndexError):                                 # one in-plane dimension only: BoundsError in the reference (:230-231)
        Fb.stream(Fb.MRI(v), lcms=Fb.MRI(np.zeros((4, 4, 4, 10), F)))
    with pytest.raises(ValueError):                                 # LCMs are [nx,ny,nz,10]
        Fb.stream(Fb.MRI(v), lcms=Fb.MRI(np.zeros((4, 4, 4, 9), F)))
    with pytest.raises(ValueError):
        Fb.stream(Fb.MRI(np.full((4, 4, 4), 100.0, F)))            # neither vectors nor angles in [-90, 90]


@pytest.mark.gpu
def test_stream_device_resident_entry_point():
    """fibers_stream_device on DEVICE volumes (peaks / amplitudes a reconstruction left on the GPU) equals the host-pointer call."""
    import ctypes as C
    import torch
    import fibers_jl_b200 as Fb
    shape = (20, 18, 10)
    vols, fs = noisy_field(shape, 2, seed=4)
    sub = Fb.draw_sublist(2, rng=8)
    ref = Fb.stream([Fb.MRI(v) for v in vols], f=[Fb.MRI(x) for x in fs], f_thresh=0.1, sublist=sub)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    d_v = [torch.from_numpy(np.ascontiguousarray(np.moveaxis(v, 3, 0).reshape(3, -1, order="F"))).to(dev) for v in vols]   # [3][nvox], x fastest
    d_f = [torch.from_numpy(np.ascontiguousarray(x.reshape(-1, order="F"))).to(dev) for x in fs]
    L = Fb._lib.lib()
    PP = C.c_void_p * 2
    h = C.c_void_p(); nstr = C.c_int64(); ntot = C.c_int64()
    Fb._lib.check(L.fibers_stream_device(PP(*[t.data_ptr() for t in d_v]), 2, *shape, PP(*[t.data_ptr() for t in d_f]), 0.1, None, 0.0, None, None,
                                          Fb._lib.ptr(sub), 2, 3, max(shape), float(np.float32(np.cos(np.deg2rad(45.0)))), 0.5, 0.2, None, 0.0,
                                          C.byref(h), C.byref(nstr), C.byref(ntot)))
    try:
        npts = np.zeros(nstr.value, np.int32); xyz = np.zeros((3, ntot.value), np.float32, order="F")
        Fb._lib.check(L.fibers_stream_fetch(h, Fb._lib.ptr(npts), Fb._lib.ptr(xyz)))
    finally:
        L.fibers_stream_free(h)
    assert nstr.value == ref.n_count > 0 and np.array_equal(npts, ref.npts)
    assert np.array_equal(xyz, np.concatenate(ref.xyz, axis=1))


def test_oracle_micro_regime_quirks():
    """Search area (src/stream.jl:268-292): unit vectors inside the unit ellipsoid, zeros outside, NaN at the centre (0 / 0), which
    the cone test cannot reject (:573-575).  Straight +x field, step 1, 5 x 5 x 1 box: the line hops to the voxel straight ahead."""
    A = SO.search_area([2, 2, 0])
    assert np.isnan(A[2, 2, 0]).all() and not A[0, 0, 0].any() and A[4, 2, 0, 0] == 1 and A[2, 0, 0, 1] == -1
    np.testing.assert_allclose(A[3, 3, 0], [np.sqrt(0.5), np.sqrt(0.5), 0], rtol=1e-6)
    v = straight_field((12, 6, 3))
    m, arr = SO.stream_work([v])
    micro = ([2, 2, 0], A, F(np.cos(np.deg2rad(10.0))))
    s = SO.new_line([5, 3, 2], np.zeros(3, F), m, arr, len_max=100, cosang_thresh=F(np.cos(np.deg2rad(20.0))), step=1.0, smooth=0.0, micro=micro)
    # every voxel of the box has |cos| = 1; the 10-degree cone around +x leaves the voxels straight ahead (dy = 0, dx > 0) and the
    # centre, whose NaN vector cannot be rejected; the centre comes first in the box's column-major order, so it wins: the line
    # advances exactly one voxel per step
    assert np.all(s[1] == 3) and np.all(s[2] == 2)
    fwd = s[0][: np.argmax(s[0] == 5.0) + 1][::-1]
    np.testing.assert_array_equal(fwd, np.arange(5.0, 5.0 + len(fwd)))


@pytest.mark.gpu
@pytest.mark.parametrize("dist", [(3, 3, 0), (2, 2, 2)])
def test_stream_gpu_micro_regime_bit_exact(dist):
    """Microscopy regime (voxel size <= 50 um): regime defaults (nsub 0, 20 degrees, step 1, no smoothing) and the box search."""
    import fibers_jl_b200 as Fb
    shape = (22, 18, 6)
    vols, _ = noisy_field(shape, 1, seed=12)
    g = np.random.default_rng(2)
    mask = np.asfortranarray((g.random(shape) < 0.92).astype(np.uint8))
    # the Python mirror always fills the three half widths with search_dist (:86); call the C ABI through it with a cubic box, and
    # the anisotropic box (through-plane 0, :153) through the oracle-equivalent keyword below
    if dist[2] == dist[0]:
        got = Fb.stream(Fb.MRI(vols[0], volres=(0.01, 0.01, 0.01)), mask=Fb.MRI(mask), search_dist=dist[0], search_ang=25)
        ref = SO.stream(vols, [np.zeros(3, F)], mask=mask, step_size=1.0, smooth_coeff=0.0, cosang_thresh=F(np.cos(np.deg2rad(np.float64(F(20))))),
                        micro_search_dist=dist, micro_search_cosang=F(np.cos(np.deg2rad(np.float64(F(25))))))
    else:
        import ctypes as C
        L = Fb._lib.lib()
        PP = C.c_void_p * 1
        h = C.c_void_p(); nstr = C.c_int64(); ntot = C.c_int64()
        sub = np.zeros((1, 3), F)
        Fb._lib.check(L.fibers_stream(PP(vols[0].ctypes.data), 1, *shape, None, 0.0, None, 0.0, Fb._lib.ptr(mask), None, Fb._lib.ptr(sub), 1, 3, max(shape),
                                      float(F(np.cos(np.deg2rad(20.0)))), 1.0, 0.0, (C.c_int32 * 3)(*dist), float(F(np.cos(np.deg2rad(25.0)))), 0,
                                      C.byref(h), C.byref(nstr), C.byref(ntot)))
        try:
            npts = np.zeros(nstr.value, np.int32); xyz = np.zeros((3, ntot.value), F, order="F")
            Fb._lib.check(L.fibers_stream_fetch(h, Fb._lib.ptr(npts), Fb._lib.ptr(xyz)))
        finally:
            L.fibers_stream_free(h)
        ends = np.cumsum(npts)
        from fibers_jl_b200.stream import Tract
        got = Tract([xyz[:, e - n:e] for e, n in zip(ends, npts)], npts)
        ref = SO.stream(vols, [np.zeros(3, F)], mask=mask, step_size=1.0, smooth_coeff=0.0, cosang_thresh=F(np.cos(np.deg2rad(20.0))),
                        micro_search_dist=dist, micro_search_cosang=F(np.cos(np.deg2rad(25.0))))
    assert got.n_count == len(ref) and got.n_count > 100
    nbad = sum(0 if np.array_equal(a, b) else 1 for a, b in zip(got.xyz, ref))
    print(f"[parity] stream micro {dist}: {got.n_count} streamlines, {int(got.npts.sum())} points, mismatching lines {nbad}")
    assert nbad == 0


@pytest.mark.gpu
def test_stream_gpu_angle_input_micro():
    """2-D orientation angles (radians) in the microscopy regime: the wrapper builds in-plane vectors (src/stream.jl:145-172), the
    search box is flat along the through-plane axis (:153); equals the oracle on those vectors."""
    import fibers_jl_b200 as Fb
    shape = (30, 24, 2)
    g = np.random.default_rng(5)
    xs, ys = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), indexing="ij")
    ang = (0.9 * np.sin(xs / 7.0) * np.cos(ys / 5.0))[..., None] + 0.05 * g.standard_normal(shape)
    ang = np.asfortranarray(np.clip(ang, -1.5, 1.5).astype(F))
    got = Fb.stream(Fb.MRI(ang, volres=(0.01, 0.01, 0.05)), search_dist=3, search_ang=30)
    vec = np.zeros(shape + (3,), F, order="F"); vec[..., 0] = np.cos(ang); vec[..., 1] = np.sin(ang); vec[ang == 0] = 0
    ref = SO.stream([vec], [np.zeros(3, F)], step_size=1.0, smooth_coeff=0.0, cosang_thresh=F(np.cos(np.deg2rad(np.float64(F(20))))),
                    micro_search_dist=(3, 3, 0), micro_search_cosang=F(np.cos(np.deg2rad(np.float64(F(30))))), len_max=30)
    assert got.n_count == len(ref) > 100 and all(np.array_equal(a, b) for a, b in zip(got.xyz, ref))


def lcm_field(shape, nvec, seed):
    """In-plane (x-y) orientation fields with holes + random local connection matrices with many sub-threshold elements."""
    g = np.random.default_rng(seed)
    nx, ny, nz = shape
    xs, ys, zs = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    vols = []
    for i in range(nvec):
        ph = g.uniform(0, 2 * np.pi, 3)
        th = 0.7 * np.sin(xs / 4.0 + ph[0]) + 0.6 * np.cos(ys / 3.0 + ph[1]) + 0.2 * np.sin(zs + ph[2]) + i * 1.3
        v = np.stack([np.cos(th), np.sin(th), np.zeros_like(th)], axis=-1)
        v[g.random(shape) < (0.04 + 0.08 * i)] = 0
        vols.append(np.asfortranarray(v.astype(F)))
    lcms = g.uniform(0, 1, shape + (10,)).astype(F)
    lcms[g.random(shape + (10,)) < 0.3] = 0
    lcms[g.random(shape) < 0.05] = 0                                  # voxels without any connection: the line ends there
    return vols, np.asfortranarray(lcms)


def test_oracle_lcm_known_answers():
    """Hand-made cases of stream_pick_by_lcm! (src/stream.jl:380-494).  2-D field in the x-y plane, two vectors per voxel:
    e1 = +x, e2 = +y.  A line running along +x enters every new voxel through its -x edge (entry edge type 3: dvox = now - next =
    (-1, 0) is column 1 ... the reference names the edge by the jump that LEAVES through it, so dvox = (-1, 0) is type 1)."""
    shape = (8, 8, 1)
    a = np.zeros(shape + (3,), F); a[..., 0] = 1
    a[7, 7, 0] = [0, 1, 0]                                            # (the reference finds the plane from the first volume: x and y must both occur)
    b = np.zeros(shape + (3,), F); b[..., 1] = 1
    m, arr = SO.stream_work([a, b])
    # (i) only element (1, 3) is set: entry 1 -> exit 3 = jump (+1, 0): the line keeps to +x whatever is drawn; flags all False
    l = np.zeros(shape + (10,), F); l[..., 2] = 1
    L = SO.lcm_work(l, 0.099, a)
    assert L[1] == [0, 1] and L[2][:, 0].tolist() == [-1, 0, 0] and L[2][:, 3].tolist() == [0, 1, 0]
    s, fl = SO.new_line([2, 4, 1], np.zeros(3, F), m, arr, 100, F(0.7), 0.5, 0.0, None, (L, lambda: F(0.5)))
    assert np.all(s[1] == 4) and s[0].max() == 8.0 and not fl.any()
    # (ii) only element (1, 4) is set: entry 1 -> exit 4 = jump (0, +1): on entering the second voxel the line turns to +y
    #      (the conventional pick would have kept +x: flag True at that point); the next voxel is entered through edge type 2
    #      (dvox = (0, -1)), which element (1, 4) does not touch: the LCM is empty and the pass ends.  The backward pass starts
    #      along the vector chosen LAST (+y, reversed), enters (2, 3) through edge 4, is sent to exit 1 (-x: flag True) and ends
    #      at the next voxel (entry 3, no element).
    l = np.zeros(shape + (10,), F); l[..., 3] = 1
    L = SO.lcm_work(l, 0.099, a)
    s, fl = SO.new_line([2, 4, 1], np.zeros(3, F), m, arr, 100, F(0.7), 0.5, 0.0, None, (L, lambda: F(0.5)))
    np.testing.assert_array_equal(s, np.array([[3, 4, 1], [2.5, 4, 1], [2, 4, 1], [2, 4, 1], [2, 3.5, 1], [2, 3, 1]], F).T)
    assert fl.tolist() == [False, True, False, False, True, False]
    s_ii = s
    # (iii) the draw decides between elements (1, 3) and (1, 4) with p = (.25, .75): u < .25 -> straight on, u >= .25 -> turn
    l = np.zeros(shape + (10,), F); l[..., 2] = 1; l[..., 3] = 3
    L = SO.lcm_work(l, 0.099, a)
    s_lo, f_lo = SO.new_line([2, 4, 1], np.zeros(3, F), m, arr, 100, F(0.7), 0.5, 0.0, None, (L, lambda: F(0.2)))
    s_hi, f_hi = SO.new_line([2, 4, 1], np.zeros(3, F), m, arr, 100, F(0.7), 0.5, 0.0, None, (L, lambda: F(0.25)))
    assert np.all(s_lo[1] == 4) and s_lo[0].max() == 8.0 and s_lo[0].min() == 1.5 and not f_lo.any()
    # (the turn of (ii); the backward pass gets one voxel further, because element (1, 3) serves entry 3 there)
    assert np.array_equal(s_hi[:, :6], s_ii) and s_hi[:, 6].tolist() == [1.5, 3, 1] and f_hi.tolist() == [False, True, False, False, True, False, False]
    # (iv) thresholding: an element below lcm_thresh is removed (compared in Float64: 0.099f0 < 0.099)
    l = np.zeros(shape + (10,), F); l[..., 2] = F(0.099)
    assert not SO.lcm_work(l, 0.099, a)[0].any()
    assert SO.lcm_work(l, float(F(0.099)), a)[0].any()
    # (v) the generator: reproducible, in [0, 1), different per line and per draw
    u = [SO.lcm_uniform(3, ln, k) for ln in range(50) for k in range(20)]
    assert min(u) >= 0 and max(u) < 1 and len(set(u)) > 990 and abs(float(np.mean(u)) - 0.5) < 0.05
    assert SO.lcm_uniform(3, 7, 9) == SO.lcm_uniform(3, 7, 9)
    assert SO.lcm_uniform(0, 0, 0) == F(_splitmix_top24(0x9E3779B97F4A7C15 + 0xD1B54A32D192ED03) * 2.0 ** -24)


def _splitmix_top24(z):
    M = 2 ** 64 - 1
    z &= M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return (z ^ (z >> 31)) >> 40


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["xy_plane", "xz_plane_mask_seed"])
def test_stream_gpu_lcm_bit_exact(case):
    """fibers_stream_lcm against the oracle: same streamlines, same method-difference flags, same draws."""
    import fibers_jl_b200 as Fb
    shape = (24, 20, 3)
    vols, lcms = lcm_field(shape, 2, seed=21)
    kw, okw = {}, {}
    if case == "xz_plane_mask_seed":                                  # through-plane = y: swap the axes of the volumes and of the vectors
        vols = [np.asfortranarray(np.swapaxes(v, 1, 2)[..., [0, 2, 1]]) for v in vols]
        lcms = np.asfortranarray(np.swapaxes(lcms, 1, 2))
        shape = (24, 3, 20)
        g = np.random.default_rng(4)
        mask = np.asfortranarray((g.random(shape) < 0.92).astype(np.uint8))
        seed = np.asfortranarray((g.random(shape) < 0.4).astype(np.uint8))
        kw.update(mask=Fb.MRI(mask), seed=Fb.MRI(seed), smooth_coeff=0.0, lcm_thresh=0.3)
        okw.update(mask=mask, seed=seed, smooth_coeff=0.0, lcm_thresh=0.3)
    sub = Fb.draw_sublist(2, rng=9)
    got = Fb.stream([Fb.MRI(v) for v in vols], lcms=Fb.MRI(lcms), lcm_seed=12345, sublist=sub, len_max=40, **kw)
    ref, rfl = SO.stream(vols, list(sub), lcms=lcms, lcm_seed=12345, len_max=40, **okw)
    assert got.n_count == len(ref) and got.n_count > 100
    assert np.array_equal(got.npts, np.array([s.shape[1] for s in ref], np.int32))
    nbad = sum(0 if np.array_equal(a, b) else 1 for a, b in zip(got.xyz, ref))
    nflag = sum(0 if np.array_equal(a[0] != 0, b) else 1 for a, b in zip(got.scalars, rfl))
    ndiff = int(sum(f.sum() for f in rfl))
    print(f"[parity] stream lcm {case}: {got.n_count} streamlines, {int(got.npts.sum())} points, {ndiff} method-difference flags, "
          f"mismatching lines {nbad}, mismatching flag rows {nflag}")
    assert nbad == 0 and nflag == 0 and ndiff > 0
    # another seed gives other lines (the draws matter)
    other = Fb.stream([Fb.MRI(v) for v in vols], lcms=Fb.MRI(lcms), lcm_seed=999, sublist=sub, len_max=40, **kw)
    assert other.n_count != got.n_count or any(not np.array_equal(a, b) for a, b in zip(other.xyz, got.xyz))


GOLD = os.path.join(ROOT, "tests", "golden")


def _golden_stream():
    g = np.load(os.path.join(GOLD, "stream_small.npz"))
    ends = np.cumsum(g["npts"])
    return g, [g["xyz"][:, e - n:e] for e, n in zip(ends, g["npts"])]


def test_oracle_reproduces_stream_fixture():
    """tests/golden/stream_small.npz (tools/make_golden.py): the fixture the pinning kit hands to the real reference."""
    g, want = _golden_stream()
    got = SO.stream(list(g["ovec"]), [np.zeros(3, F)], f=list(g["f"]), f_thresh=float(g["f_thresh"]), mask=g["mask"])
    assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))


@pytest.mark.gpu
def test_stream_gpu_matches_golden_fixture():
    import fibers_jl_b200 as Fb
    g, want = _golden_stream()
    tr = Fb.stream([Fb.MRI(np.asfortranarray(v)) for v in g["ovec"]], f=[Fb.MRI(np.asfortranarray(x)) for x in g["f"]], f_thresh=float(g["f_thresh"]),
                   mask=Fb.MRI(np.asfortranarray(g["mask"])), nsub=0)
    assert tr.n_count == len(want) and np.array_equal(tr.npts, g["npts"])
    assert all(np.array_equal(a, b) for a, b in zip(tr.xyz, want))
