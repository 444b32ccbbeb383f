"""End-to-end fibers_gqi_rec on PAGEABLE numpy arrays (the bounce-ring path) for a few copy-thread counts and chunk sizes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, bench
import fibers_jl_b200 as F
shape = (145, 174, 145); nvox = int(np.prod(shape))
bval, bvec = bench.make_tables(); nvol = 288
dev = torch.device("cuda", 0)
dwi = bench.synth_dwi_device(torch, nvox, bval, bvec, 1, dev)
mask = torch.ones(nvox, dtype=torch.uint8, device=dev)
pg = bench.HostSubject(torch, shape, nvol, False, (dwi, mask))
del dwi
L = F._lib.lib()
V = np.asfortranarray(F.sphere_642.vertices); Fc = np.asfortranarray(F.sphere_642.faces); bv = np.asfortranarray(bvec)
def call():
    F._lib.check(L.fibers_gqi_rec(pg.dwi_ptr, 0, pg.mask_ptr, *shape, nvol, F._lib.ptr(bval), F._lib.ptr(bv), F._lib.ptr(V), V.shape[0],
                                  F._lib.ptr(Fc), Fc.shape[0], 1.25, *pg.gqi_out(), None, 1))
for threads in (4, 8, 16, 32):
    for chunk in (1 << 16, 1 << 17, 1 << 18):
        os.environ["FIBERS_CUDA_COPY_THREADS"] = str(threads); os.environ["FIBERS_CUDA_CHUNK_VOXELS"] = str(chunk)
        L.fibers_cuda_release_cache()
        med, mean = bench.timed_calls(call, 1, 3)
        print(f"copy threads {threads:2d} chunk {chunk:7d}: {med*1e3:7.1f} ms/call  {nvox/med:.3e} voxels/s", flush=True)
