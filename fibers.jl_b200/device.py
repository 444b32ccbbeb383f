"""Device-resident entry points of libfibers_cuda.so (plans + slab calls on raw device pointers).

Pointers are plain integers (e.g. `tensor.data_ptr()`), streams are cudaStream_t handles as
integers (e.g. `torch.cuda.current_stream().cuda_stream`); this module itself does not import
torch.  Used by bench.py (kernel-only timing), the slab/batch drivers and the GPU tests.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .odf import ODF, sphere_642


class Plan:
    """GPU analogue of GQIwork / DSIwork / DTIwork / ADCwork on one device."""

    def __init__(self, kind: str, device: int, bval, bvec=None, odf_dirs: ODF = sphere_642, sigma: float = 1.25,
                 hann_width: int = 32):
        L = _lib.lib()
        self.kind = kind
        self.device = device
        self.nvol = int(np.asarray(bval).shape[0])
        bval = np.ascontiguousarray(bval, np.float32)
        bv = None if bvec is None else np.asfortranarray(bvec, np.float32)
        h = C.c_void_p()
        if kind == "dti":
            rc = L.fibers_dti_plan_create(C.byref(h), device, self.nvol, _lib.ptr(bval), _lib.ptr(bv))
        elif kind == "adc":
            rc = L.fibers_adc_plan_create(C.byref(h), device, self.nvol, _lib.ptr(bval))
        elif kind in ("gqi", "dsi"):
            V = np.asfortranarray(odf_dirs.vertices, np.float32)
            F = np.asfortranarray(odf_dirs.faces, np.int32)
            self.nvert = odf_dirs.nvert
            if kind == "gqi":
                rc = L.fibers_gqi_plan_create(C.byref(h), device, self.nvol, _lib.ptr(bval), _lib.ptr(bv), _lib.ptr(V),
                                              V.shape[0], _lib.ptr(F), F.shape[0], float(np.float32(sigma)))
            else:
                rc = L.fibers_dsi_plan_create(C.byref(h), device, self.nvol, _lib.ptr(bval), _lib.ptr(bv), _lib.ptr(V),
                                              V.shape[0], _lib.ptr(F), F.shape[0], int(hann_width))
        else:
            raise ValueError(kind)
        _lib.check(rc)
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:       # (module globals are None during interpreter shutdown)
            _lib.lib().fibers_plan_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def kernel(self) -> str:
        return {1: "simt", 2: "tc"}.get(_lib.lib().fibers_plan_kernel(self._h), "?")

    def matrix(self) -> np.ndarray:
        L = _lib.lib()
        rows = L.fibers_plan_matrix(self._h, None, 0)
        out = np.zeros((rows, self.nvol), np.float32)
        assert L.fibers_plan_matrix(self._h, _lib.ptr(out), out.size) == rows
        return out

    # --- slab calls (all arguments are device pointers as ints; pitches in elements) ----------
    def recon(self, d_dwi, dwi_pitch, d_mask, nvox, out_pitch, d_odf, d_peak, d_qa, d_stats, d_pdf=0, d_peak_idx=0,
              finalize=True, stream=0):
        _lib.check(_lib.lib().fibers_recon_device(self._h, d_dwi, dwi_pitch, d_mask, nvox, out_pitch, d_pdf or None,
                                                  d_odf, d_peak[0], d_peak[1], d_peak[2], d_qa[0], d_qa[1], d_qa[2],
                                                  d_peak_idx or None, d_stats, 1 if finalize else 0, stream or None))

    def dti_fit(self, d_dwi, dwi_pitch, d_mask, nvox, out_pitch, d_out10, d_valid=0, stream=0):
        _lib.check(_lib.lib().fibers_dti_fit_device(self._h, d_dwi, dwi_pitch, d_mask, nvox, out_pitch, *d_out10,
                                                    d_valid or None, stream or None))

    def adc_fit(self, d_dwi, dwi_pitch, d_mask, nvox, d_adc, d_s0, stream=0):
        _lib.check(_lib.lib().fibers_adc_fit_device(self._h, d_dwi, dwi_pitch, d_mask, nvox, d_adc, d_s0, stream or None))


def stats_init(d_stats, stream=0):
    _lib.check(_lib.lib().fibers_stats_init_device(d_stats, stream or None))


def qa_scale(d_qa, nvox, d_stats=0, odfmax=0.0, stream=0):
    _lib.check(_lib.lib().fibers_qa_scale_device(d_qa[0], d_qa[1], d_qa[2], nvox, d_stats or None, float(odfmax),
                                                 stream or None))


def decode_max(encoded: int) -> float:
    return float(_lib.lib().fibers_stats_decode_max(int(encoded)))


def launch_count() -> int:
    return int(_lib.lib().fibers_cuda_launch_count())


def set_kernel(name: str):
    _lib.check(_lib.lib().fibers_cuda_set_kernel({"auto": 0, "simt": 1, "tc": 2}[name]))


def set_devices(devs):
    arr = (C.c_int * len(devs))(*devs)
    _lib.check(_lib.lib().fibers_cuda_set_devices(arr, len(devs)))
