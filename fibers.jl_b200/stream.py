"""Host-side mirror of the reference's `stream` (src/stream.jl:730-790): same keyword arguments, same defaults,
same output order; one blocking call into libfibers_cuda.so (fibers_stream) plus the fetch of the result.

Covers what has a deterministic answer -- orientation vectors, no local connection matrices, in the macroscopic and in
the microscopy regime (voxel size <= 50 um: regime-dependent defaults and the box search of stream_micro_new_point!,
src/stream.jl:84-95, :547-617); 2-D orientation ANGLES (:145-172) are turned into in-plane vectors here, as the
StreamWork constructor does.  `lcms` (random sampling from the connection matrix, :394-492) is not on the GPU path and
raises.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .mri import MRI


@dataclass
class Tract:                       # the streamline fields of the reference's Tract (src/trk.jl:37-41) that stream() fills
    xyz: list                      # [3, npts] float32 per streamline, 1-based voxel coordinates
    npts: np.ndarray               # int32 [nstr]
    sublist: np.ndarray = field(default=None, repr=False)   # the sub-voxel offsets that were used (the reference keeps them in StreamWork)

    ref: dict = field(default=None, repr=False)             # volsize / volres / vox2ras of the volume the header is built from (Tract{T}(ref::MRI))
    scalars: list = field(default=None, repr=False)         # trk_read: [n_scalars, npts] float32 per streamline
    properties: np.ndarray = field(default=None, repr=False)  # trk_read: [n_properties, nstr] float32
    header: dict = field(default=None, repr=False)          # trk_read: the remaining header fields of the reference's Tract (src/trk.jl:11-35)

    @property
    def n_count(self):
        return len(self.xyz)


def trk_write(tr: "Tract", outfile: str) -> bool:
    """trk_write(tr, outfile) -- src/trk.jl:433-495; header as Tract{T}(ref::MRI) builds it (:88-145).  Returns True on error."""
    ref = tr.ref or {}
    xyz = np.concatenate(tr.xyz, axis=1) if len(tr.xyz) else np.zeros((3, 0), np.float32)
    xyz = np.asfortranarray(xyz, dtype=np.float32)
    vs = np.ascontiguousarray(ref.get("volsize", [1, 1, 1]), np.int32)
    vr = np.ascontiguousarray(ref.get("volres", [1, 1, 1]), np.float32)
    M = np.ascontiguousarray(ref.get("vox2ras0", np.eye(4)), np.float32)
    npts = np.ascontiguousarray(tr.npts, np.int32)
    sc = pr = None
    ns = npr = 0
    if tr.scalars is not None and len(tr.scalars) and tr.scalars[0].shape[0] > 0:
        sc = np.asfortranarray(np.concatenate(tr.scalars, axis=1), dtype=np.float32); ns = sc.shape[0]
    if tr.properties is not None and np.size(tr.properties):
        pr = np.asfortranarray(np.asarray(tr.properties, np.float32).reshape(-1, len(npts))); npr = pr.shape[0]
    _lib.check(_lib.lib().fibers_trk_write_ex(outfile.encode(), _lib.ptr(vs), _lib.ptr(vr), _lib.ptr(M), int(len(npts)), _lib.ptr(npts), _lib.ptr(xyz),
                                              ns, _lib.ptr(sc) if ns else None, npr, _lib.ptr(pr) if npr else None))
    return False


class TrkInfo(C.Structure):            # fibers_trk_info (include/fibers_cuda.h)
    _fields_ = [("dim", C.c_int32 * 3), ("voxel_size", C.c_float * 3), ("origin", C.c_float * 3), ("n_scalars", C.c_int32),
                ("n_properties", C.c_int32), ("vox_to_ras", C.c_float * 16), ("voxel_order", C.c_char * 4),
                ("voxel_order_original", C.c_char * 4), ("image_orientation_patient", C.c_float * 6), ("n_count", C.c_int32),
                ("version", C.c_int32), ("hdr_size", C.c_int32), ("total_points", C.c_int64)]


def trk_read(infile: str) -> "Tract":
    """trk_read(infile) -- src/trk.jl:358-425: points come back as xyz ./ voxel_size .- .5 (0-based voxel coordinates);
    scalars per point and properties per streamline as stored."""
    L = _lib.lib()
    info = TrkInfo()
    _lib.check(L.fibers_trk_read_info(infile.encode(), C.byref(info)))
    n, tot, ns, npr = info.n_count, info.total_points, info.n_scalars, info.n_properties
    npts = np.zeros(n, np.int32)
    xyz = np.zeros((3, tot), np.float32, order="F")
    sc = np.zeros((ns, tot), np.float32, order="F")
    pr = np.zeros((npr, n), np.float32, order="F")
    _lib.check(L.fibers_trk_read_data(infile.encode(), C.byref(info), _lib.ptr(npts), _lib.ptr(xyz), _lib.ptr(sc) if ns else None,
                                      _lib.ptr(pr) if npr else None))
    cuts = np.cumsum(npts)[:-1]
    vs = np.array(info.voxel_size[:], np.float32)
    M = np.array(info.vox_to_ras[:], np.float32).reshape(4, 4)
    hdr = {"dim": np.array(info.dim[:], np.int16), "voxel_size": vs, "origin": np.array(info.origin[:], np.float32), "n_scalars": ns,
           "n_properties": npr, "vox_to_ras": M, "voxel_order": bytes(info.voxel_order), "voxel_order_original": bytes(info.voxel_order_original),
           "image_orientation_patient": np.array(info.image_orientation_patient[:], np.float32), "n_count": n, "version": info.version,
           "hdr_size": info.hdr_size}
    return Tract(xyz=[np.asfortranarray(a) for a in np.split(xyz, cuts, axis=1)] if n else [], npts=npts,
                 ref={"volsize": [int(d) for d in info.dim], "volres": vs, "vox2ras0": M},
                 scalars=[np.asfortranarray(a) for a in np.split(sc, cuts, axis=1)] if n else [], properties=pr, header=hdr)


def _vol(m):
    return m.vol if isinstance(m, MRI) else np.asarray(m)


def draw_sublist(nsub: int, rng=None) -> np.ndarray:
    """Sub-voxel sampling offsets, src/stream.jl:177-183: `T.(rand(Uniform(-.5+eps(), .5-eps()), 3))` per sample, or one
    zero offset when nsub == 0."""
    if nsub <= 0:
        return np.zeros((1, 3), np.float32)
    g = np.random.default_rng(rng)
    e = np.finfo(np.float64).eps
    return g.uniform(-0.5 + e, 0.5 - e, size=(nsub, 3)).astype(np.float32)


def stream(ovec, *, odf=None, f=None, f_thresh=0.03, fa=None, fa_thresh=0.1, mask=None, seed=None, nsub=None, len_min=3,
           len_max=None, ang_thresh=None, step_size=None, smooth_coeff=None, search_dist=15, search_ang=10, lcms=None,
           lcm_thresh=0.099, verbose=False, sublist=None, rng=None, device=0, timing=None, lcm_seed=None) -> Tract:
    """stream(ovec; odf, f, f_thresh, fa, fa_thresh, mask, seed, nsub, len_min, len_max, ang_thresh, step_size,
    smooth_coeff, search_dist, search_ang, lcms, lcm_thresh, verbose) -- reference: src/stream.jl:730.

    `ovec`: MRI or list of MRI, each [nx,ny,nz,3]; `f`: MRI or list of MRI [nx,ny,nz].  Extra keywords: `sublist`
    ([nsub,3] float32; drawn like the reference draws them when omitted, with `rng` as the seed), `lcm_seed` (seed of the
    counter-based generator behind the LCM draws; drawn from `rng` when omitted) and `device`."""
    ovecs = list(ovec) if isinstance(ovec, (list, tuple)) else [ovec]
    fs = None if f is None else (list(f) if isinstance(f, (list, tuple)) else [f])
    res = ovecs[0].header.get("volres") if isinstance(ovecs[0], MRI) else None
    domicro = res is not None and min(res) <= 0.05                   # microscopy regime (src/stream.jl:84)
    vols = []
    thrudim = None
    ovec0_given = None
    for o in ovecs:
        v = np.asfortranarray(_vol(o), dtype=np.float32)
        if v.ndim == 3 or (v.ndim == 4 and v.shape[3] == 1):
            # 2-D orientation ANGLES (src/stream.jl:145-172): in-plane unit vectors (cos, sin); the through-plane axis is the one
            # with the largest voxel size.  Done here, like the StreamWork constructor does it; the library only sees vectors.
            ang = v.reshape(v.shape[:3], order="F")
            r3 = res if res is not None else (1.0, 1.0, 1.0)
            thrudim = int(np.argmax(r3))
            s1, s2 = [d for d in range(3) if d != thrudim]
            vec = np.zeros(ang.shape + (3,), np.float32, order="F")
            eps = np.finfo(np.float32).eps
            if -np.pi / 2 - eps <= ang.min() and ang.max() <= np.pi / 2 + eps:          # radians
                vec[..., s1] = np.cos(ang); vec[..., s2] = np.sin(ang)
            elif -90 <= ang.min() and ang.max() <= 90:                                  # degrees
                vec[..., s1] = np.cos(np.deg2rad(ang.astype(np.float64))).astype(np.float32)
                vec[..., s2] = np.sin(np.deg2rad(ang.astype(np.float64))).astype(np.float32)
            else:
                raise ValueError("Input orientations should be 3D vectors or angles in [-90, 90]")
            if mask is None:                                                            # mask = any(x -> x != 0, vol) (:107-112)
                vec[ang == 0] = 0
            v = vec
        if v.ndim != 4 or v.shape[3] != 3:
            raise _lib.FibersCudaError(1, "stream: orientation volumes must be [nx,ny,nz,3] vectors or [nx,ny,nz] angles")
        if ovec0_given is None:
            ovec0_given = v
        vols.append(v)
    nx, ny, nz = vols[0].shape[:3]
    if any(v.shape[:3] != (nx, ny, nz) for v in vols):
        raise ValueError("stream: orientation volumes differ in size")
    nvec = len(vols)
    if fs is not None and len(fs) != nvec:
        raise ValueError("stream: one amplitude volume per orientation volume is required")
    fvols = None if fs is None else [np.asfortranarray(_vol(x), dtype=np.float32).reshape((nx, ny, nz), order="F") for x in fs]
    fav = None if fa is None else np.asfortranarray(np.asarray(_vol(fa), dtype=np.float32).reshape((nx, ny, nz, -1), order="F")[..., 0])
    mk = None
    if mask is not None:
        mv = _vol(mask)
        mk = np.asfortranarray((mv.reshape((nx, ny, nz, -1), order="F")[..., 0] > 0).astype(np.uint8))
    sd = None
    if seed is not None:
        sv = _vol(seed)
        if mask is not None and tuple(sv.shape) != tuple(_vol(mask).shape):
            raise ValueError(f"Dimension mismatch between seed mask {tuple(sv.shape)} and brain mask {tuple(_vol(mask).shape)}")
        sd = np.asfortranarray((sv.reshape((nx, ny, nz, -1), order="F")[..., 0] > 0).astype(np.uint8))
    # defaults that depend on the regime (src/stream.jl:91-95)
    if nsub is None:
        nsub = 0 if domicro else 3
    if sublist is None:
        sublist = draw_sublist(int(nsub), rng)
    sub = np.ascontiguousarray(sublist, dtype=np.float32).reshape(-1, 3)
    if len_max is None:
        len_max = max(nx, ny, nz)                                   # maximum(ovec.volsize)
    ang = (20 if domicro else 45) if ang_thresh is None else ang_thresh
    step = (1 if domicro else 0.5) if step_size is None else step_size
    smooth = (0 if domicro else 0.2) if smooth_coeff is None else smooth_coeff
    msd = None; mcos = 0.0
    if domicro:
        sd3 = [int(search_dist)] * 3                                 # fill(Int(search_dist), 3) (:86)
        if thrudim is not None:
            sd3[thrudim] = 0                                         # 2-D data: no search through the plane (:153)
        msd = (C.c_int32 * 3)(*sd3)
        mcos = float(np.float32(np.cos(np.deg2rad(np.float64(np.float32(search_ang))))))   # cosd(T(search_ang)) (:306)
    cos_thresh = np.float32(np.cos(np.deg2rad(np.float64(np.float32(ang)))))    # cosd(T(ang_thresh))

    if lcms is not None:
        if domicro:
            raise ValueError("stream: local connection matrices are only defined for the macroscopic regime")
        lv = np.asfortranarray(_vol(lcms), dtype=np.float32)
        if lv.shape != (nx, ny, nz, 10):
            raise ValueError(f"stream: lcms must be [nx,ny,nz,10], got {lv.shape}")
        # through-plane dimension = the component of the first orientation volume that is zero everywhere (:224-226)
        thru = [d for d in range(3) if not np.any(ovec0_given[..., d])]
        strdims = [d for d in range(3) if d not in thru]
        if len(strdims) < 2:
            raise IndexError("stream: the first orientation volume has fewer than two non-zero components (BoundsError in the reference, :230-231)")
        if lcm_seed is None:                                         # the reference draws from Julia's RNG: a fresh seed per call unless one is given
            lcm_seed = int(np.random.default_rng(rng).integers(0, 2 ** 63))
    L = _lib.lib(); _lib.require_device()
    PP = C.c_void_p * nvec
    ov_ptrs = PP(*[v.ctypes.data for v in vols])
    f_ptrs = None if fvols is None else PP(*[v.ctypes.data for v in fvols])
    import time
    handle = C.c_void_p(); nstr = C.c_int64(); ntot = C.c_int64()
    t0 = time.perf_counter()
    scal = None
    if lcms is not None:
        _lib.check(L.fibers_stream_lcm(ov_ptrs, nvec, nx, ny, nz, f_ptrs, float(f_thresh), _lib.ptr(fav), float(fa_thresh), _lib.ptr(mk),
                                       _lib.ptr(sd), _lib.ptr(sub), int(sub.shape[0]), int(len_min), int(len_max), float(step), float(smooth),
                                       _lib.ptr(lv), float(lcm_thresh), int(strdims[0]), int(strdims[1]), int(lcm_seed) & (2 ** 64 - 1),
                                       int(device), C.byref(handle), C.byref(nstr), C.byref(ntot)))
    else:
        _lib.check(L.fibers_stream(ov_ptrs, nvec, nx, ny, nz, f_ptrs, float(f_thresh), _lib.ptr(fav), float(fa_thresh), _lib.ptr(mk),
                                   _lib.ptr(sd), _lib.ptr(sub), int(sub.shape[0]), int(len_min), int(len_max), float(cos_thresh),
                                   float(step), float(smooth), msd, mcos, int(device), C.byref(handle), C.byref(nstr), C.byref(ntot)))
    t1 = time.perf_counter()
    try:
        npts = np.zeros(nstr.value, np.int32)
        xyz = np.zeros((3, ntot.value), np.float32, order="F")
        if handle.value:
            _lib.check(L.fibers_stream_fetch(handle, _lib.ptr(npts), _lib.ptr(xyz)))
            if lcms is not None:
                scal = np.zeros((1, ntot.value), np.float32, order="F")
                _lib.check(L.fibers_stream_fetch_scalars(handle, _lib.ptr(scal)))
        elif lcms is not None:
            scal = np.zeros((1, 0), np.float32, order="F")
    finally:
        L.fibers_stream_free(handle)
    if timing is not None:                     # (bench aid) seconds in the tracking call and in the fetch of the points
        timing["call_s"] = t1 - t0; timing["fetch_s"] = time.perf_counter() - t1
    ends = np.cumsum(npts, dtype=np.int64)
    lines = [xyz[:, e - n:e] for e, n in zip(ends, npts)]
    href = mask if isinstance(mask, MRI) else (ovecs[0] if isinstance(ovecs[0], MRI) else None)      # Tract{Float32}(mask) (:783)
    ref = None
    if href is not None:
        M0 = np.asarray(href.header.get("vox2ras0", np.eye(4)), np.float32)
        ref = dict(volsize=[nx, ny, nz], vox2ras0=M0,
                   volres=href.header.get("volres", np.sqrt((M0[:3, :3].astype(np.float64) ** 2).sum(0)).tolist()))
    tr = Tract(lines, npts, sub, ref)
    if scal is not None:
        tr.scalars = [scal[:, e - n:e] for e, n in zip(ends, npts)]
    return tr
