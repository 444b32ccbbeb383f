"""RUMBA-SD timing: rumba_rec through the C ABI on a brain-masked HCP-shaped volume (sphere_724, niter iterations),
per-iteration cost, and the numpy oracle on a small sub-volume as the CPU reference point.
    python tools/gpu/bench_rumba.py [--shape 145,174,145] [--niter 50] [--fill 0.25]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch, bench
import fibers_jl_b200 as F
from fibers_jl_b200 import phantom

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="145,174,145"); ap.add_argument("--niter", type=int, default=50); ap.add_argument("--fill", type=float, default=0.25)
ap.add_argument("--cpu-shape", default="24,24,12")
a = ap.parse_args()
shape = tuple(int(x) for x in a.shape.split(",")); nvox = int(np.prod(shape))
bval, bvec = bench.make_tables()
dev = torch.device("cuda", 0)
dwi = bench.synth_dwi_device(torch, nvox, bval, bvec, 11, dev)
h = np.zeros(shape + (bval.shape[0],), np.float32, order="F")
h.reshape(-1, bval.shape[0], order="F")[...] = dwi.T.cpu().numpy()
del dwi; torch.cuda.empty_cache()
mask = phantom.ellipsoid_mask(shape, a.fill)
nmask = int(mask.sum())
D, M = F.MRI(h, bval, bvec), F.MRI(mask)
times = {}
for n in (0, a.niter):
    F.rumba_rec(D, M, niter=min(n, 2))
    t = time.perf_counter(); r = F.rumba_rec(D, M, niter=n); times[n] = time.perf_counter() - t
per_iter = (times[a.niter] - times[0]) / a.niter
ndir, ncomp = int((bval != bval.min()).sum()) + 1, 364
flops = 3 * 2.0 * ndir * ncomp * nmask
out = {"kernel": "rumba_rec", "shape": shape, "mask_voxels": nmask, "niter": a.niter, "s_total": times[a.niter], "s_setup_and_output": times[0],
       "ms_per_iteration": per_iter * 1e3, "gemm_TFLOPs_fp32": flops / per_iter / 1e12, "s_600_iterations_extrapolated": times[0] + 600 * per_iter,
       "snr_mean": r.snr_mean}
# CPU reference point: the numpy float32 oracle (BLAS sgemm + vectorised numpy) on a small volume, scaled per mask voxel and iteration
import rumba_oracle as R
cs = tuple(int(x) for x in a.cpu_shape.split(","))
ph = phantom.gqi_phantom(cs, seed=3, mask_fill=0.8)
t = time.perf_counter(); R.rumba_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], np.asarray(F.sphere_724.vertices), niter=5, dtype=np.float32); dt = time.perf_counter() - t
cpu_per = dt / 5 / int((ph["mask"] > 0).sum())
out["cpu_oracle_us_per_voxel_iteration"] = cpu_per * 1e6
out["gpu_us_per_voxel_iteration"] = per_iter / nmask * 1e6
out["cpu_cores"] = len(os.sched_getaffinity(0))
print(json.dumps(out))
