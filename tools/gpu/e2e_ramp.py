"""A/B of the host pipeline's ramp-up (FIBERS_CUDA_RAMP) and chunk size on the host-pointer entry point fibers_gqi_rec (pinned buffers)."""
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
h_dwi = torch.empty((nvol, nvox), dtype=torch.float32, pin_memory=True); h_dwi.copy_(dwi); torch.cuda.synchronize(); del dwi
h_mask = torch.ones(nvox, dtype=torch.uint8, pin_memory=True)
h_odf = torch.empty((321, nvox), dtype=torch.float32, pin_memory=True)
h_peak = [torch.empty((3, nvox), dtype=torch.float32, pin_memory=True) for _ in range(3)]
h_qa = [torch.empty(nvox, dtype=torch.float32, pin_memory=True) for _ in range(3)]
L = F._lib.lib()
V = np.asfortranarray(F.sphere_642.vertices); Fc = np.asfortranarray(F.sphere_642.faces); bv = np.asfortranarray(bvec)
def call():
    F._lib.check(L.fibers_gqi_rec(h_dwi.data_ptr(), 0, h_mask.data_ptr(), *shape, nvol, F._lib.ptr(bval), F._lib.ptr(bv), F._lib.ptr(V),
                                  V.shape[0], F._lib.ptr(Fc), Fc.shape[0], 1.25, h_odf.data_ptr(), *[p.data_ptr() for p in h_peak],
                                  *[q.data_ptr() for q in h_qa], None, 1))
cases = [("ramp off, chunk 2^18", {"FIBERS_CUDA_RAMP": "0"}), ("ramp on,  chunk 2^18", {}), ("ramp on,  chunk 2^19", {"FIBERS_CUDA_CHUNK_VOXELS": str(1 << 19)}),
         ("ramp on,  chunk 2^17", {"FIBERS_CUDA_CHUNK_VOXELS": str(1 << 17)})]
for rep in range(2):
    for label, env in cases:
        for k in ("FIBERS_CUDA_RAMP", "FIBERS_CUDA_CHUNK_VOXELS"): os.environ.pop(k, None)
        L.fibers_cuda_release_cache()
        os.environ.update(env)
        call(); call(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            t = time.perf_counter(); call(); ts.append(time.perf_counter() - t)
        dt = float(np.median(ts))
        print(f"{label:22s} {dt*1e3:8.1f} ms/call  {nvox/dt:.3e} voxels/s", flush=True)
