"""fibers.jl_b200 -- B200-native voxel-wise diffusion reconstruction behind the Fibers.jl API.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + the C ABI of
libfibers_cuda.so), the host-side mirror of the reference's function signatures (`recon.py`),
the MRI / ODF containers those signatures take, and synthetic phantoms.  Import as
`fibers_jl_b200` (see the loader module at the repo root: the directory name contains a dot).
"""
from . import _lib, device, batch
from ._lib import FibersCudaError, device_count
from .mri import MRI
from .odf import ODF, sphere_362, sphere_642, sphere_724
from .stream import Tract, stream, draw_sublist, trk_write, trk_read
from .io import mri_read, mri_write, mri_read_bfiles, mri_filename
from .recon import DTI, GQI, DSI, RUMBASD, dti_write, gqi_write, dsi_write, rumba_write, rumba_rec, st_eigen, st_recon, adc_fit, dti_fit, gqi_rec, dsi_rec, dti_gqi_fit, dti_gqi_fit_batch

__all__ = ["MRI", "ODF", "sphere_362", "sphere_642", "sphere_724", "DTI", "GQI", "DSI",
           "RUMBASD", "rumba_rec", "st_eigen", "st_recon", "adc_fit", "dti_fit", "gqi_rec", "dsi_rec", "dti_gqi_fit", "dti_gqi_fit_batch", "Tract", "stream", "draw_sublist", "trk_write", "trk_read", "dti_write", "gqi_write", "dsi_write", "rumba_write", "mri_read", "mri_write", "mri_read_bfiles", "mri_filename", "FibersCudaError", "device_count"]
