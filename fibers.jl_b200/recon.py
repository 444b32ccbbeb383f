"""Host-side mirror of the reference's reconstruction API (same names, argument meaning and
error behaviour), each function one blocking call into libfibers_cuda.so:

    adc_fit(dwi, mask)                                    reference: src/dti.jl:164
    dti_fit(dwi, mask)                                    reference: src/dti.jl:221
    gqi_rec(dwi, mask, odf_dirs=sphere_642, sigma=1.25)   reference: src/gqi.jl:109
    dsi_rec(dwi, mask, odf_dirs=sphere_642, hann_width=32)  reference: src/dsi.jl:171

This is exactly what julia/FibersCUDA.jl does with `ccall`; Python stands in for Julia because
Julia is not installed in this image.  There is no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .mri import MRI
from .odf import ODF, sphere_362, sphere_642, sphere_724


@dataclass
class DTI:                     # reference: src/dti.jl:11-22 (field order = file-name suffixes)
    s0: MRI
    eigval1: MRI
    eigval2: MRI
    eigval3: MRI
    eigvec1: MRI
    eigvec2: MRI
    eigvec3: MRI
    rd: MRI
    md: MRI
    fa: MRI
    valid: np.ndarray | None = None    # extra (test aid): voxels that took a fit branch


@dataclass
class GQI:                     # reference: src/gqi.jl:10-14
    odf: MRI
    peak: list
    qa: list
    peak_idx: np.ndarray | None = None   # extra (test aid): 0-based vertex index or -1


@dataclass
class DSI:                     # reference: src/dsi.jl:10-15
    pdf: MRI
    odf: MRI
    peak: list
    qa: list
    peak_idx: np.ndarray | None = None


@dataclass
class RUMBASD:                 # reference: src/rusd.jl:11-20
    fodf: MRI
    fgm: MRI
    fcsf: MRI
    peak: list
    gfa: MRI
    var: MRI
    snr_mean: float
    snr_std: float
    peak_idx: np.ndarray | None = None   # extra (test aid): 0-based vertex index or -1, [nx,ny,nz,5]


def _write_fields(obj, basename: str, names) -> None:
    """The loop shared by dti_write / gqi_write / dsi_write / rumba_write: MRI fields go to <basename>_<field>.nii.gz, vectors of
    MRI to <basename>_<field><i>.nii.gz (i from 1), anything else (RUMBA-SD's SNR estimates) to <basename>_<field>.txt."""
    from .io import mri_write
    for name in names:
        val = getattr(obj, name)
        if isinstance(val, MRI):
            mri_write(val, f"{basename}_{name}.nii.gz")
        elif isinstance(val, (list, tuple)):
            for i, m in enumerate(val):
                mri_write(m, f"{basename}_{name}{i + 1}.nii.gz")
        else:                                              # writedlm(fname, value, ' ')
            with open(f"{basename}_{name}.txt", "w") as f:
                f.write(" ".join(str(x) for x in np.atleast_1d(val).ravel()) + "\n")


def dti_write(dti: "DTI", basename: str) -> None:
    """dti_write(dti, basename) -- src/dti.jl:344-349: one NIfTI volume per field of the structure."""
    _write_fields(dti, basename, ("s0", "eigval1", "eigval2", "eigval3", "eigvec1", "eigvec2", "eigvec3", "rd", "md", "fa"))


def gqi_write(gqi: "GQI", basename: str) -> None:
    """gqi_write(gqi, basename) -- src/gqi.jl:210-226."""
    _write_fields(gqi, basename, ("odf", "peak", "qa"))


def dsi_write(dsi: "DSI", basename: str) -> None:
    """dsi_write(dsi, basename) -- src/dsi.jl:279-295."""
    _write_fields(dsi, basename, ("pdf", "odf", "peak", "qa"))


def rumba_write(rumba: "RUMBASD", basename: str) -> None:
    """rumba_write(rumba, basename) -- src/rusd.jl:645-664 (the SNR estimates go to text files)."""
    _write_fields(rumba, basename, ("fodf", "fgm", "fcsf", "peak", "gfa", "var", "snr_mean", "snr_std"))


def _mask_u8(mask: MRI, shape):
    m = np.asarray(mask.vol)
    if m.ndim == 4 and m.shape[3] == 1:          # tutorial masks are [nx,ny,nz,1]
        m = m[..., 0]
    if tuple(m.shape) != tuple(shape):
        raise _lib.FibersCudaError(1, f"mask size {m.shape} does not match dwi size {shape}")
    return np.asfortranarray(m != 0).astype(np.uint8, order="F")


def _check_tables(dwi: MRI, need_bvec: bool):
    if dwi.bval is None or dwi.bval.size == 0:
        raise RuntimeError("Missing b-value table from input DWI structure")      # src/dti.jl:166-168
    if need_bvec and (dwi.bvec is None or dwi.bvec.size == 0):
        raise RuntimeError("Missing gradient table from input DWI structure")     # src/dti.jl:227-229
    nvol = dwi.vol.shape[3]
    if dwi.bval.shape[0] != nvol or (need_bvec and dwi.bvec.shape[0] != nvol):
        raise _lib.FibersCudaError(1, "b-table length does not match the number of volumes")


def adc_fit(dwi: MRI, mask: MRI, ngpu: int = 1):
    """Fit the apparent diffusion coefficient; returns (adc, s0) as MRI structures."""
    _check_tables(dwi, False)
    if dwi.vol.dtype != np.float32:
        raise TypeError("adc_fit requires a Float32 DWI volume (reference method signature, src/dti.jl:197)")
    L = _lib.lib(); _lib.require_device()
    nx, ny, nz, nvol = dwi.vol.shape
    m = _mask_u8(mask, (nx, ny, nz))
    adc, s0 = MRI.like(mask, 1), MRI.like(mask, 1)
    vol = np.asfortranarray(dwi.vol)
    _lib.check(L.fibers_adc_fit(_lib.ptr(vol), _lib.ptr(m), nx, ny, nz, nvol, _lib.ptr(dwi.bval),
                                _lib.ptr(adc.vol), _lib.ptr(s0.vol), ngpu))
    return adc, s0


def dti_fit(dwi: MRI, mask: MRI, ngpu: int = 1) -> DTI:
    """Fit tensors to DWIs and return a `DTI` structure."""
    _check_tables(dwi, True)
    if dwi.vol.dtype != np.float32:
        raise TypeError("dti_fit requires a Float32 DWI volume (reference method signature, src/dti.jl:286)")
    L = _lib.lib(); _lib.require_device()
    nx, ny, nz, nvol = dwi.vol.shape
    m = _mask_u8(mask, (nx, ny, nz))
    outs = [MRI.like(mask, n) for n in (1, 1, 1, 1, 3, 3, 3, 1, 1, 1)]
    valid = np.zeros((nx, ny, nz), np.uint8, order="F")
    vol = np.asfortranarray(dwi.vol)
    bvec = np.asfortranarray(dwi.bvec, np.float32)
    _lib.check(L.fibers_dti_fit(_lib.ptr(vol), _lib.ptr(m), nx, ny, nz, nvol, _lib.ptr(dwi.bval), _lib.ptr(bvec),
                                *[_lib.ptr(o.vol) for o in outs], _lib.ptr(valid), ngpu))
    return DTI(*outs, valid=valid.astype(bool))


def _recon(kind, dwi: MRI, mask: MRI, odf_dirs: ODF, param, ngpu: int, want_odf: bool = True):
    _check_tables(dwi, True)
    L = _lib.lib(); _lib.require_device()
    nx, ny, nz, nvol = dwi.vol.shape
    m = _mask_u8(mask, (nx, ny, nz))
    vol = np.asfortranarray(dwi.vol)
    code = _lib.dtype_code(vol.dtype)
    M = odf_dirs.nvert
    odf = MRI.like(mask, M) if want_odf else None        # odf=NULL: peaks / QA only, the ODF never crosses PCIe
    peak = [MRI.like(mask, 3) for _ in range(3)]
    qa = [MRI.like(mask, 1) for _ in range(3)]
    idx = np.zeros((nx, ny, nz, 3), np.int16, order="F")
    bvec = np.asfortranarray(dwi.bvec, np.float32)
    V = np.asfortranarray(odf_dirs.vertices, np.float32)
    F = np.asfortranarray(odf_dirs.faces, np.int32)            # Matrix{Integer} -> Matrix{Int32}
    common = (_lib.ptr(vol), code, _lib.ptr(m), nx, ny, nz, nvol, _lib.ptr(dwi.bval), _lib.ptr(bvec),
              _lib.ptr(V), V.shape[0], _lib.ptr(F), F.shape[0])
    tail = (_lib.ptr(odf.vol) if want_odf else None, *[_lib.ptr(p.vol) for p in peak], *[_lib.ptr(q.vol) for q in qa], _lib.ptr(idx), ngpu)
    if kind == "gqi":
        _lib.check(L.fibers_gqi_rec(*common, float(np.float32(param)), *tail))
        return GQI(odf, peak, qa, idx)
    pdf = MRI.like(mask, nvol)
    _lib.check(L.fibers_dsi_rec(*common, int(param), _lib.ptr(pdf.vol), *tail))
    return DSI(pdf, odf, peak, qa, idx)


def gqi_rec(dwi: MRI, mask: MRI, odf_dirs: ODF = sphere_642, sigma: float = 1.25, ngpu: int = 1, want_odf: bool = True) -> GQI:
    """Generalized q-sampling imaging reconstruction; returns a `GQI` structure.
    `want_odf=False` (extension): `.odf` is None, only peaks and QA come back (what `stream` consumes)."""
    return _recon("gqi", dwi, mask, odf_dirs, sigma, ngpu, want_odf)


def dti_gqi_fit(dwi: MRI, mask: MRI, odf_dirs: ODF = sphere_642, sigma: float = 1.25, ngpu: int = 1):
    """`dti_fit(dwi, mask)` and `gqi_rec(dwi, mask, odf_dirs, sigma)` in one pass over the DWI volume (each z-slab
    chunk crosses PCIe once and feeds both kernels); returns `(DTI, GQI)`, bit-identical to the two separate calls
    (src/dti.jl:221, src/gqi.jl:109)."""
    _check_tables(dwi, True)
    if dwi.vol.dtype != np.float32:
        raise TypeError("dti_fit requires a Float32 DWI volume (reference method signature, src/dti.jl:286)")
    L = _lib.lib(); _lib.require_device()
    nx, ny, nz, nvol = dwi.vol.shape
    m = _mask_u8(mask, (nx, ny, nz))
    vol = np.asfortranarray(dwi.vol)
    bvec = np.asfortranarray(dwi.bvec, np.float32)
    V = np.asfortranarray(odf_dirs.vertices, np.float32)
    F = np.asfortranarray(odf_dirs.faces, np.int32)
    douts = [MRI.like(mask, n) for n in (1, 1, 1, 1, 3, 3, 3, 1, 1, 1)]
    odf = MRI.like(mask, odf_dirs.nvert)
    peak = [MRI.like(mask, 3) for _ in range(3)]
    qa = [MRI.like(mask, 1) for _ in range(3)]
    _lib.check(L.fibers_dti_gqi_fit(_lib.ptr(vol), _lib.ptr(m), nx, ny, nz, nvol, _lib.ptr(dwi.bval), _lib.ptr(bvec),
                                    *[_lib.ptr(o.vol) for o in douts], _lib.ptr(V), V.shape[0], _lib.ptr(F), F.shape[0],
                                    float(np.float32(sigma)), _lib.ptr(odf.vol), *[_lib.ptr(p.vol) for p in peak],
                                    *[_lib.ptr(q.vol) for q in qa], ngpu))
    return DTI(*douts), GQI(odf, peak, qa, None)


def dsi_rec(dwi: MRI, mask: MRI, odf_dirs: ODF = sphere_642, hann_width: int = 32, ngpu: int = 1, want_odf: bool = True) -> DSI:
    """Diffusion spectrum imaging reconstruction; returns a `DSI` structure."""
    return _recon("dsi", dwi, mask, odf_dirs, hann_width, ngpu, want_odf)


def dti_gqi_fit_batch(dwis, masks, odf_dirs: ODF = sphere_642, sigma: float = 1.25, ngpu: int = 1, want_dti: bool = True,
                      want_odf: bool = True):
    """`dti_fit` + `gqi_rec` of a batch of subjects that share one protocol and one volume shape (BASELINE cfg4) in
    ONE library call: subjects are queued over `ngpu` devices and every device overlaps the transfers of the next
    subject with the kernels of the current one.  Returns a list of `(DTI | None, GQI)`, bit-identical to calling
    `dti_gqi_fit` per subject (src/dti.jl:221, src/gqi.jl:109)."""
    import ctypes as C
    nsub = len(dwis)
    if nsub != len(masks):
        raise _lib.FibersCudaError(1, "dwis and masks must have the same length")
    if nsub == 0:
        return []
    for d in dwis:
        _check_tables(d, True)
        if d.vol.dtype != np.float32:
            raise TypeError("dti_fit requires a Float32 DWI volume (reference method signature, src/dti.jl:286)")
        if d.vol.shape != dwis[0].vol.shape or not np.array_equal(d.bval, dwis[0].bval) or not np.array_equal(d.bvec, dwis[0].bvec):
            raise _lib.FibersCudaError(1, "all subjects of a batch must share the volume shape and the b-table")
    L = _lib.lib(); _lib.require_device()
    nx, ny, nz, nvol = dwis[0].vol.shape
    ms = [_mask_u8(m, (nx, ny, nz)) for m in masks]
    vols = [np.asfortranarray(d.vol) for d in dwis]
    bvec = np.asfortranarray(dwis[0].bvec, np.float32)
    V = np.asfortranarray(odf_dirs.vertices, np.float32)
    Fc = np.asfortranarray(odf_dirs.faces, np.int32)
    res, dptr, gptr = [], [], []
    for m in masks:
        douts = [MRI.like(m, n) for n in (1, 1, 1, 1, 3, 3, 3, 1, 1, 1)] if want_dti else None
        odf = MRI.like(m, odf_dirs.nvert) if want_odf else None
        peak = [MRI.like(m, 3) for _ in range(3)]
        qa = [MRI.like(m, 1) for _ in range(3)]
        if want_dti:
            dptr += [o.vol.ctypes.data for o in douts]
        gptr += [odf.vol.ctypes.data if want_odf else None] + [p.vol.ctypes.data for p in peak] + [q.vol.ctypes.data for q in qa]
        res.append((DTI(*douts) if want_dti else None, GQI(odf, peak, qa, None)))
    parr = lambda xs: (C.c_void_p * len(xs))(*xs)
    _lib.check(L.fibers_dti_gqi_fit_batch(nsub, parr([v.ctypes.data for v in vols]), parr([m.ctypes.data for m in ms]),
                                          nx, ny, nz, nvol, _lib.ptr(dwis[0].bval), _lib.ptr(bvec),
                                          parr(dptr) if want_dti else None, _lib.ptr(V), V.shape[0], _lib.ptr(Fc), Fc.shape[0],
                                          float(np.float32(sigma)), parr(gptr), ngpu))
    return res


def rumba_rec(dwi: MRI, mask: MRI, odf_dirs: ODF = sphere_724, niter: int = 600, lambda_para: float = 1.7e-3,
              lambda_perp: float = 0.2e-3, lambda_csf: float = 3.0e-3, lambda_gm: float = 0.8e-4, ncoils: int = 1,
              coil_combine: str = "SMF-SENSE", ipat_factor: int = 1, use_tv: bool = True, device: int = 0) -> RUMBASD:
    """Robust and unbiased model-based spherical deconvolution (RUMBA-SD); returns a `RUMBASD` structure.
    Same arguments, defaults and error behaviour as the reference's `rumba_rec` (src/rusd.jl:419)."""
    _check_tables(dwi, True)
    if coil_combine not in ("SMF-SENSE", "SoS-GRAPPA"):
        raise RuntimeError("Unknown coil combine mode " + coil_combine)          # src/rusd.jl:432-434
    if ipat_factor < 1:
        raise RuntimeError("iPAT factor must be a positive integer")              # :436-438
    if dwi.vol.dtype != np.float32:
        raise TypeError("rumba_rec requires a Float32 DWI volume (reference method signature, src/rusd.jl:419)")
    nv2 = odf_dirs.vertices.shape[0]
    # the reference picks the neighbourhood by comparing with its own spheres (:479-483); any other mesh is an UndefVarError there
    if nv2 in (724, 642):
        ang_neig = 12.5
    elif nv2 == 362:
        ang_neig = 16.0
    else:
        raise RuntimeError("rumba_rec: odf_dirs must be sphere_724, sphere_642 or sphere_362 (angular neighbourhood undefined otherwise)")
    L = _lib.lib(); _lib.require_device()
    nx, ny, nz, nvol = dwi.vol.shape
    m = np.asarray(mask.vol)
    if m.ndim == 4 and m.shape[3] == 1:
        m = m[..., 0]
    if tuple(m.shape) != (nx, ny, nz):
        raise _lib.FibersCudaError(1, f"mask size {m.shape} does not match dwi size {(nx, ny, nz)}")
    mpos = np.asfortranarray(m > 0).astype(np.uint8, order="F")
    many = np.asfortranarray(m != 0).astype(np.uint8, order="F")
    vol = np.asfortranarray(dwi.vol)
    bvec = np.asfortranarray(dwi.bvec, np.float32)
    V = np.asfortranarray(odf_dirs.vertices, np.float32)
    nvert = nv2 // 2
    fodf = MRI.like(mask, nvert)
    fgm, fcsf, gfa, var = (MRI.like(mask, 1) for _ in range(4))
    peak = [MRI.like(mask, 3) for _ in range(5)]
    idx = np.zeros((nx, ny, nz, 5), np.int16, order="F")
    import ctypes as C
    sm, ss = C.c_float(0), C.c_float(0)
    _lib.check(L.fibers_rumba_rec(_lib.ptr(vol), _lib.ptr(mpos), _lib.ptr(many), nx, ny, nz, nvol, _lib.ptr(dwi.bval), _lib.ptr(bvec),
                                  _lib.ptr(V), nv2, float(ang_neig), int(niter), float(np.float32(lambda_para)), float(np.float32(lambda_perp)),
                                  float(np.float32(lambda_csf)), float(np.float32(lambda_gm)), int(ncoils), 1 if coil_combine == "SoS-GRAPPA" else 0,
                                  int(ipat_factor), 1 if use_tv else 0, _lib.ptr(fodf.vol), _lib.ptr(fgm.vol), _lib.ptr(fcsf.vol),
                                  *[_lib.ptr(p.vol) for p in peak], _lib.ptr(gfa.vol), _lib.ptr(var.vol), C.byref(sm), C.byref(ss),
                                  _lib.ptr(idx), int(device)))
    return RUMBASD(fodf, fgm, fcsf, peak, gfa, var, float(sm.value), float(ss.value), idx)


def st_eigen(Sxx, Sxy, Sxz, Syy, Syz, Szz, device: int = 0):
    """Eigen-decomposition of a structure-tensor field; returns `(eigvec [nx,ny,nz,3,3], eigval [nx,ny,nz,3])` exactly as the
    reference's `st_eigen` (src/structens.jl:13-34): eigenvalues ascending, `eigvec[x,y,z,:,k]` the k-th eigenvector."""
    arrs = [np.asfortranarray(a) for a in (Sxx, Sxy, Sxz, Syy, Syz, Szz)]
    if any(a.dtype != np.float32 for a in arrs):
        raise TypeError("st_eigen: Float32 volumes only on the GPU path")
    if any(a.ndim != 3 or a.shape != arrs[0].shape for a in arrs):
        raise _lib.FibersCudaError(1, "st_eigen: the six tensor components must be 3-D arrays of one size")
    L = _lib.lib(); _lib.require_device()
    nx, ny, nz = arrs[0].shape
    evec = np.zeros((nx, ny, nz, 3, 3), np.float32, order="F"); evals = np.zeros((nx, ny, nz, 3), np.float32, order="F")
    _lib.check(L.fibers_st_eigen(*[_lib.ptr(a) for a in arrs], nx, ny, nz, _lib.ptr(evec), _lib.ptr(evals), int(device)))
    return evec, evals


def st_recon(vol, sigma: float, rho: float, device: int = 0):
    """Structure-tensor reconstruction of a 3-D image (src/structens.jl:40-88): Gaussian pre-smoothing `sigma`, Scharr gradients,
    Gaussian tensor smoothing `rho`, eigen-decomposition.  Returns `(eigvec, eigval)` like `st_eigen`."""
    v = np.asfortranarray(vol)
    if v.dtype != np.float32 or v.ndim != 3:
        raise TypeError("st_recon: a 3-D Float32 volume is required on the GPU path")
    L = _lib.lib(); _lib.require_device()
    nx, ny, nz = v.shape
    evec = np.zeros((nx, ny, nz, 3, 3), np.float32, order="F"); evals = np.zeros((nx, ny, nz, 3), np.float32, order="F")
    _lib.check(L.fibers_st_recon(_lib.ptr(v), nx, ny, nz, float(sigma), float(rho), _lib.ptr(evec), _lib.ptr(evals), int(device)))
    return evec, evals
