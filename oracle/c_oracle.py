"""ctypes driver for the C/OpenMP port (oracle/fibers_oracle.c).  TEST INFRASTRUCTURE / CPU
BASELINE ONLY -- see the header of fibers_oracle.py.  The matrices the reference precomputes
(pinv(A), sinc matrix, DSI tables) come from the numpy oracle; the C code runs the voxel loops.

NOTE for the GPU box: the .so is compiled with -march=native where it is BUILT; `load()` rebuilds
it when the cached binary was built on a different CPU model."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import fibers_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libfibers_oracle.so")
_lib = None


def _cpu_tag():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def build(force=False):
    tag_file = SO + ".cpu"
    tag = _cpu_tag()
    stale = (not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, "fibers_oracle.c"))
             or not os.path.exists(tag_file) or open(tag_file).read() != tag)
    if force or stale:
        subprocess.run(["make", "-C", HERE, "-B"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        open(tag_file, "w").write(tag)
    return SO


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(load().oracle_max_threads())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _parr(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def dti_fit(dwi, mask, bval, bvec, nthreads=0):
    L = load()
    nx, ny, nz, nvol = dwi.shape
    ib0, A, pA = O.dti_design(bval, bvec, np.float32)
    dwi = np.asfortranarray(dwi, np.float32); m = np.asfortranarray(np.asarray(mask) != 0).astype(np.uint8, order="F")
    names = ["s0", "eigval1", "eigval2", "eigval3", "eigvec1", "eigvec2", "eigvec3", "rd", "md", "fa"]
    nfr = [1, 1, 1, 1, 3, 3, 3, 1, 1, 1]
    outs = [np.zeros((nx, ny, nz) + ((n,) if n > 1 else ()), np.float32, order="F") for n in nfr]
    valid = np.zeros((nx, ny, nz), np.uint8, order="F")
    A = np.ascontiguousarray(A); pA = np.ascontiguousarray(pA); ib = np.ascontiguousarray(ib0.astype(np.uint8))
    L.oracle_linfit(_p(dwi), _p(m), nx, ny, nz, nvol, 7, _p(A), _p(pA), _p(ib), _parr(outs), _p(valid), nthreads)
    r = dict(zip(names, outs)); r["valid"] = valid.astype(bool)
    return r


def adc_fit(dwi, mask, bval, nthreads=0):
    L = load()
    nx, ny, nz, nvol = dwi.shape
    ib0, A, pA = O.adc_design(bval, np.float32)
    dwi = np.asfortranarray(dwi, np.float32); m = np.asfortranarray(np.asarray(mask) != 0).astype(np.uint8, order="F")
    outs = [np.zeros((nx, ny, nz), np.float32, order="F") for _ in range(2)]
    A = np.ascontiguousarray(A); pA = np.ascontiguousarray(pA); ib = np.ascontiguousarray(ib0.astype(np.uint8))
    L.oracle_linfit(_p(dwi), _p(m), nx, ny, nz, nvol, 2, _p(A), _p(pA), _p(ib), _parr(outs), None, nthreads)
    return outs[0], outs[1]


def _recon_outputs(nx, ny, nz, M):
    odf = np.zeros((nx, ny, nz, M), np.float32, order="F")
    peak = [np.zeros((nx, ny, nz, 3), np.float32, order="F") for _ in range(3)]
    qa = [np.zeros((nx, ny, nz), np.float32, order="F") for _ in range(3)]
    idx = np.zeros((nx, ny, nz, 3), np.int16, order="F")
    return odf, peak, qa, idx


class GqiSetup:
    """GQIwork equivalent, reusable across timed calls."""
    def __init__(self, bval, bvec, vertices, faces, sigma=1.25):
        self.M = vertices.shape[0] // 2
        self.A = np.ascontiguousarray(O.gqi_matrix(bval, bvec, vertices, sigma, np.float32))
        self.faces = np.ascontiguousarray((O.fold_faces(faces, self.M) - 1).astype(np.int32))
        self.vert = np.ascontiguousarray(vertices[:self.M], np.float32)


def gqi_rec(dwi, mask, bval=None, bvec=None, vertices=None, faces=None, sigma=1.25, nthreads=0, setup=None, outputs=None):
    L = load()
    nx, ny, nz, nvol = dwi.shape
    S = setup or GqiSetup(bval, bvec, vertices, faces, sigma)
    dwi = np.asfortranarray(dwi, np.float32); m = np.asfortranarray(np.asarray(mask) != 0).astype(np.uint8, order="F")
    odf, peak, qa, idx = outputs or _recon_outputs(nx, ny, nz, S.M)
    with np.errstate(all="ignore"):
        L.oracle_gqi_rec(_p(dwi), _p(m), nx, ny, nz, nvol, _p(S.A), S.M, _p(S.faces), S.faces.shape[0], _p(S.vert),
                         _p(odf), _parr(peak), _parr(qa), _p(idx), nthreads)
    return dict(odf=odf, peak=peak, qa=qa, peak_idx=idx)


def dsi_rec(dwi, mask, bval, bvec, vertices, faces, hann_width=32, nthreads=0):
    L = load()
    nx, ny, nz, nvol = dwi.shape
    M = vertices.shape[0] // 2
    W = O.dsi_work(bval, bvec, vertices, hann_width, np.float32)
    nfft = W["nfft"]
    sub = W["iq_sub"] - 1
    iq_lin = np.ascontiguousarray((sub[:, 0] + nfft * (sub[:, 1] + nfft * sub[:, 2])).astype(np.int32))
    H = np.ascontiguousarray(W["H"].reshape(-1, order="F"), np.float32)       # x fastest
    coords = np.ascontiguousarray(np.transpose(W["coords"], (2, 1, 0)), np.float32)   # [M][nrad][3]
    qr2 = np.ascontiguousarray(W["qr2"], np.float32)
    F = np.ascontiguousarray((O.fold_faces(faces, M) - 1).astype(np.int32))
    V = np.ascontiguousarray(vertices[:M], np.float32)
    dwi = np.asfortranarray(dwi, np.float32); m = np.asfortranarray(np.asarray(mask) != 0).astype(np.uint8, order="F")
    odf, peak, qa, idx = _recon_outputs(nx, ny, nz, M)
    pdf = np.zeros((nx, ny, nz, nvol), np.float32, order="F")
    L.oracle_dsi_rec.argtypes = None
    with np.errstate(all="ignore"):
        L.oracle_dsi_rec(_p(dwi), _p(m), nx, ny, nz, nvol, nfft, _p(iq_lin), _p(H), _p(coords), _p(qr2),
                         int(qr2.shape[0]), C.c_float(float(W["dqr"])), M, _p(F), F.shape[0], _p(V), _p(pdf), _p(odf),
                         _parr(peak), _parr(qa), _p(idx), nthreads)
    return dict(pdf=pdf, odf=odf, peak=peak, qa=qa, peak_idx=idx)
