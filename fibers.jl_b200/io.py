"""Host-side mirror of the reference's volume I/O (src/mri.jl): `mri_read` (:611-720), `mri_write` (:1695-1937),
`mri_read_bfiles` (:2179-2240), `mri_filename` (:520-560).  The byte-level work (NIfTI-1 / MGH headers, byte order, scaling,
gzip) is done by libfibers_cuda.so (fibers_mri_read_info / _read_data / _write); this module only allocates the arrays and
fills the `MRI` container, as the Julia wrapper would.

`mri_read(..., pin=True)` page-locks the volume it returns (fibers_cuda_host_register), so that the reconstruction entry
points DMA straight from it instead of going through the bounce ring.  Bruker directories are not read."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .mri import MRI

_NP = {_lib.F32: np.float32, _lib.F64: np.float64, _lib.I16: np.int16, _lib.U16: np.uint16, _lib.I32: np.int32, _lib.U8: np.uint8,
       _lib.I8: np.int8, _lib.U32: np.uint32, _lib.I64: np.int64}
_CODE = {np.dtype(v): k for k, v in _NP.items()}


class MriInfo(C.Structure):            # fibers_mri_info (include/fibers_cuda.h)
    _fields_ = [("format", C.c_int32), ("gz", C.c_int32), ("ndim", C.c_int32), ("dim", C.c_int32 * 4), ("dtype", C.c_int32),
                ("bswap", C.c_int32), ("sform_code", C.c_int32), ("qform_code", C.c_int32), ("data_offset", C.c_int64),
                ("vox2ras0", C.c_float * 16), ("sform", C.c_float * 16), ("qform", C.c_float * 16), ("pixdim", C.c_float * 8),
                ("volres", C.c_float * 3), ("tr", C.c_float), ("flip_angle", C.c_float), ("te", C.c_float), ("ti", C.c_float),
                ("scl_slope", C.c_float), ("scl_inter", C.c_float)]


EXTLIST = ["mgh", "mgz", "nii", "nii.gz"]


def mri_filename(fstring: str, checkdisk: bool = True):
    """(fname, fstem, fext) -- src/mri.jl:520-560."""
    fname = fstem = fext = ""
    idot = fstring.rfind(".")
    if idot < 0 and checkdisk:
        for ext in EXTLIST:
            if os.path.isfile(fstring + "." + ext):
                fname, fstem, fext = fstring + "." + ext, fstring, ext
    elif idot >= 0:
        ext = fstring[idot + 1:].lower()
        if ext == "gz":
            j = fstring.rfind(".", 0, idot)
            if j >= 0:
                idot = j
                ext = fstring[idot + 1:].lower()
        if ext in EXTLIST:
            fname, fstem, fext = fstring, fstring[:idot], ext
    return fname, fstem, fext


def _readdlm(path):
    with open(path) as f:
        rows = [[np.float32(x) for x in line.replace(",", " ").split()] for line in f if line.strip()]
    w = max(len(r) for r in rows)
    if any(len(r) != w for r in rows):
        raise ValueError(f"File {path} contains non-numeric entries")
    return np.array(rows, np.float32).reshape(len(rows), w)


def mri_read_bfiles(infile1: str, infile2: str):
    """b-value vector [n] and gradient table [n, 3] from two text files given in any order (src/mri.jl:2179-2240)."""
    tab = []
    for f in (infile1, infile2):
        if not os.path.isfile(f):
            raise FileNotFoundError("Could not open " + f)
        tab.append(_readdlm(f))
    ival, ivec = (0, 1) if tab[0].size < tab[1].size else (1, 0)
    names = (infile1, infile2)
    if tab[ival].shape[1] != 1:
        if tab[ival].shape[0] != 1:
            raise ValueError("Wrong format in table " + names[ival] + " (should be single column or row)")
        tab[ival] = tab[ival].T
    if tab[ivec].shape[1] != 3:
        if tab[ivec].shape[0] != 3:
            raise ValueError("Wrong format in table " + names[ivec] + " (should be three columns or rows)")
        tab[ivec] = tab[ivec].T
    if tab[0].shape[0] != tab[1].shape[0]:
        raise ValueError(f"Dimension mismatch between tables in {infile1} {tab[0].shape} and {infile2} {tab[1].shape}")
    return np.ascontiguousarray(tab[ival][:, 0]), np.asfortranarray(tab[ivec])


def mri_read(infile: str, headeronly: bool = False, permutedata: bool = False, pin: bool = False) -> MRI:
    """mri_read(infile; headeronly, permutedata) -- src/mri.jl:611-720 (MGH and NIfTI branches)."""
    if os.path.isdir(infile):
        raise _lib.FibersCudaError(1, "mri_read: Bruker scan directories are not read by this library")
    fname, fstem, fext = mri_filename(infile)
    if not fname:
        raise ValueError("Cannot determine format of " + infile)
    L = _lib.lib()
    info = MriInfo()
    _lib.check(L.fibers_mri_read_info(fname.encode(), C.byref(info)))
    nx, ny, nz, nfr = (int(x) for x in info.dim)
    shape = (nx, ny, nz, nfr) if info.ndim >= 4 else (nx, ny, nz)
    dt = _NP[info.dtype]
    if headeronly:
        vol = np.zeros((0,) * len(shape), dt, order="F")
    else:
        vol = np.zeros(shape, dt, order="F")
        if pin:
            _lib.check(L.fibers_cuda_host_register(_lib.ptr(vol), vol.nbytes))
        _lib.check(L.fibers_mri_read_data(fname.encode(), C.byref(info), _lib.ptr(vol), vol.nbytes))
    M = np.array(info.vox2ras0, np.float32).reshape(4, 4)
    hdr = dict(fspec=fname, vox2ras0=M, vox2ras1=_vox2ras_0to1(M), volsize=[nx, ny, nz], nframes=nfr, volres=[float(x) for x in info.volres],
               tr=float(info.tr), flip_angle=float(info.flip_angle), te=float(info.te), ti=float(info.ti), ispermuted=False,
               width=nx, height=ny, depth=nz)
    if info.format == 1:
        hdr["niftihdr"] = dict(scl_slope=float(info.scl_slope), scl_inter=float(info.scl_inter), sform_code=int(info.sform_code),
                               qform_code=int(info.qform_code), sform=np.array(info.sform, np.float32).reshape(4, 4),
                               qform=np.array(info.qform, np.float32).reshape(4, 4), pixdim=[float(x) for x in info.pixdim],
                               do_bswap=bool(info.bswap), vox_offset=int(info.data_offset))
    mri = MRI(None, **hdr)
    mri.vol = vol
    # optional DWI tables (:679-705)
    bfile = next((fstem + e for e in (".bvals", ".bval") if os.path.isfile(fstem + e)), "")
    gfile = next((fstem + e for e in (".bvecs", ".bvec") if os.path.isfile(fstem + e)), "")
    if bfile and gfile:
        b, g = mri_read_bfiles(bfile, gfile)
        if len(b) == nfr:
            with np.errstate(invalid="ignore", divide="ignore"):
                g = (g / np.sqrt(np.sum(g * g, axis=1, keepdims=True, dtype=np.float32))).astype(np.float32)   # normalise the gradient vectors (:701-703)
            g[np.isnan(g)] = 0
            mri.bval, mri.bvec = b, np.asfortranarray(g)
    if permutedata:                                                  # (:714-719)
        if not headeronly:
            mri.vol = np.asfortranarray(np.swapaxes(mri.vol, 0, 1))
        vs, vr = mri.header["volsize"], mri.header["volres"]
        mri.header["volsize"] = [vs[1], vs[0], vs[2]]; mri.header["volres"] = [vr[1], vr[0], vr[2]]
        mri.header["ispermuted"] = True
    return mri


def _vox2ras_0to1(M0):
    """src/mri.jl:327-345: vox2ras for 1-based indices."""
    Q = np.zeros((4, 4), M0.dtype); Q[:3, 3] = 1
    return (np.linalg.inv(np.linalg.inv(M0.astype(np.float64)) + Q)).astype(M0.dtype)


def mri_write(mri: MRI, outfile: str, datatype=None) -> bool:
    """mri_write(mri, outfile, datatype=eltype(mri.vol)) -- src/mri.jl:1695-1937.  Returns True if an error occurred (the
    reference's convention); also writes <stem>.bvals / <stem>.bvecs when the tables are present."""
    if mri.vol is None or mri.vol.size == 0:
        raise ValueError("Input structure has empty vol field")
    fname, fstem, fext = mri_filename(outfile, False)
    if not fname:
        raise ValueError("Cannot determine format of " + outfile)
    vol = mri.vol
    if mri.header.get("ispermuted"):
        vol = np.swapaxes(vol, 0, 1)
    vol = np.asfortranarray(vol)
    if vol.dtype not in _CODE:
        raise ValueError(f"Data type {vol.dtype} not supported")
    shp = list(vol.shape) + [1] * (4 - vol.ndim)
    dim = (C.c_int32 * 4)(*shp[:4])
    M = np.ascontiguousarray(mri.header.get("vox2ras0", np.eye(4)), dtype=np.float32)
    vr = mri.header.get("volres")
    if vr is not None and mri.header.get("ispermuted"):
        vr = [vr[1], vr[0], vr[2]]
    vres = None if vr is None else np.ascontiguousarray(vr, np.float32)
    nh = mri.header.get("niftihdr", {})
    out_dt = -1 if datatype is None else _CODE[np.dtype(datatype)]
    L = _lib.lib()
    rc = L.fibers_mri_write(fname.encode(), _lib.ptr(vol), _CODE[vol.dtype], dim, _lib.ptr(M), _lib.ptr(vres),
                            float(mri.header.get("tr", 0)), float(mri.header.get("flip_angle", 0)), float(mri.header.get("te", 0)),
                            float(mri.header.get("ti", 0)), float(nh.get("scl_slope", 0)), float(nh.get("scl_inter", 0)), out_dt)
    _lib.check(rc)
    if mri.bval is not None and len(mri.bval):                       # writedlm(bfile, mri.bval, ' ') (:1923-1931)
        np.savetxt(fstem + ".bvals", np.asarray(mri.bval, np.float32), fmt="%s")
    if mri.bvec is not None and len(mri.bvec):
        np.savetxt(fstem + ".bvecs", np.asarray(mri.bvec, np.float32), fmt="%s", delimiter=" ")
    return False
