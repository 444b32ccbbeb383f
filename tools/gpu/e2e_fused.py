"""End-to-end (host pointers, pinned buffers) time of DTI + GQI on one cfg2-shaped subject: the two reference-facing
calls one after the other against the fused fibers_dti_gqi_fit (one H2D of the DWI volume)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
import fibers_jl_b200 as F

shape = (145, 174, 145)
nvox = int(np.prod(shape))
bval, bvec = bench.make_tables()
nvol = bval.shape[0]
dev = torch.device("cuda", 0)
dwi = bench.synth_dwi_device(torch, nvox, bval, bvec, 1, dev)
h_dwi = torch.empty((nvol, nvox), dtype=torch.float32, pin_memory=True); h_dwi.copy_(dwi); del dwi
h_mask = torch.ones(nvox, dtype=torch.uint8).pin_memory()
pin = lambda n: torch.empty((n, nvox), dtype=torch.float32, pin_memory=True)
dti = [pin(n) for n in (1, 1, 1, 1, 3, 3, 3, 1, 1, 1)]
odf = pin(321); peak = [pin(3) for _ in range(3)]; qa = [pin(1) for _ in range(3)]
L = F._lib.lib()
V = np.asfortranarray(F.sphere_642.vertices, np.float32); Fc = np.asfortranarray(F.sphere_642.faces, np.int32)
bv = np.asfortranarray(bvec, np.float32)
P = F._lib.ptr

def run_dti():
    F._lib.check(L.fibers_dti_fit(h_dwi.data_ptr(), h_mask.data_ptr(), *shape, nvol, P(bval), P(bv), *[o.data_ptr() for o in dti], None, 1))
def run_gqi():
    F._lib.check(L.fibers_gqi_rec(h_dwi.data_ptr(), 0, h_mask.data_ptr(), *shape, nvol, P(bval), P(bv), P(V), V.shape[0], P(Fc), Fc.shape[0],
                                  1.25, odf.data_ptr(), *[p.data_ptr() for p in peak], *[q.data_ptr() for q in qa], None, 1))
def run_fused():
    F._lib.check(L.fibers_dti_gqi_fit(h_dwi.data_ptr(), h_mask.data_ptr(), *shape, nvol, P(bval), P(bv), *[o.data_ptr() for o in dti],
                                      P(V), V.shape[0], P(Fc), Fc.shape[0], 1.25, odf.data_ptr(), *[p.data_ptr() for p in peak],
                                      *[q.data_ptr() for q in qa], 1))
def timeit(fn, n=5):
    fn(); fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return float(np.median(ts)) * 1e3
res = {"dti_fit_ms": timeit(run_dti), "gqi_rec_ms": timeit(run_gqi)}
res["separate_ms"] = timeit(lambda: (run_dti(), run_gqi()))
res["fused_dti_gqi_fit_ms"] = timeit(run_fused)
res["voxels"] = nvox
res["note"] = "cfg2-shaped subject 145x174x145x288, mask == 1, pinned host buffers, 1 GPU; median of 5 calls after 2 warm-ups"
print(json.dumps(res))
