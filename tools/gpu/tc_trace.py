"""Debug aid: run the tensor-core kernel once with FIBERS_TC_TRACE and print the per-role clock
trace of CTA 0 (cycles relative to the first MMA of the first tile)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["FIBERS_TC_TRACE"] = "/tmp/tc_trace.bin"
import torch
import bench
import fibers_jl_b200 as F
from fibers_jl_b200 import device as D

shape = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "145,174,145").split(","))
nvox = int(np.prod(shape))
bval, bvec = bench.make_tables()
if os.environ.get("TC_TRACE_CFG5"):                      # 8 b0 + 120 directions at b = 4000 (128 volumes)
    from fibers_jl_b200 import phantom
    bval, bvec = phantom.shells_table(8, [(4000.0, 120)])
dev = torch.device("cuda", 0)
mask = torch.ones(nvox, dtype=torch.uint8, device=dev)
pitch = (nvox + 63) // 64 * 64
dwi = bench.synth_dwi_device(torch, nvox, bval, bvec, 1, dev, pitch=pitch)
odf = torch.empty((321, pitch), dtype=torch.float32, device=dev)
peak = [torch.empty((3, pitch), dtype=torch.float32, device=dev) for _ in range(3)]
qa = [torch.empty(nvox, dtype=torch.float32, device=dev) for _ in range(3)]
stats = torch.zeros(2, dtype=torch.int32, device=dev)
D.set_kernel("tc")
plan = D.Plan("gqi", 0, bval, bvec, F.sphere_642, 1.25)
for _ in range(2):
    plan.recon(dwi.data_ptr(), pitch, mask.data_ptr(), nvox, pitch, odf.data_ptr(), [p.data_ptr() for p in peak],
               [q.data_ptr() for q in qa], stats.data_ptr(), finalize=True, stream=0)
torch.cuda.synchronize()
raw = np.fromfile("/tmp/tc_trace.bin", dtype=np.int64)
t = raw[:512].reshape(16, 32)
span = raw[512:].reshape(-1, 2)
cyc = span[127]
span = span[:127]
span = span[span[:, 1] > 0]
if len(span):
    dur = (span[:, 1] - span[:, 0]) / 1e3
    print(f"per-cluster kernel span (us): n={len(dur)} min {dur.min():.0f} median {np.median(dur):.0f} max {dur.max():.0f}; "
          f"start spread {(span[:, 0].max() - span[:, 0].min()) / 1e3:.0f} us, end spread {(span[:, 1].max() - span[:, 1].min()) / 1e3:.0f} us")
    print(f"cluster 0: {cyc[1] - cyc[0]} SM cycles in {(span[0, 1] - span[0, 0]) / 1e3:.0f} us -> effective SM clock {(cyc[1] - cyc[0]) / (span[0, 1] - span[0, 0]) * 1e3:.0f} MHz")
    print("slowest clusters:", np.argsort(-dur)[:8].tolist(), " fastest:", np.argsort(dur)[:8].tolist())
t0 = t[0, 0]
names = {0: "mma:start", 1: "mma:done", 2: "epi:d_full", 3: "epi:tmem_read_done", 4: "epi:settle(prev)_done+bar", 7: "epi:outputs(prev)_done",
         5: "epi:scan_done+bar", 9: "conv:tile_start", 10: "conv:first_loads_issued", 11: "conv:a_empty_ok", 12: "conv:tile_end",
         13: "tma:tile_start", 14: "tma:tile_end"}
for it in range(8):
    ev = sorted((int(t[it, k] - t0), names[k]) for k in names if t[it, k])
    print(f"tile {it}: " + "  ".join(f"{n}@{c}" for c, n in ev))
    print(f"         mma wait b_full {int(t[it,15])} cyc, wait a_full {int(t[it,16])} cyc; drain chunk starts (rel. d_full) "
          + " ".join(str(int(t[it, k] - t[it, 2])) for k in range(17, 24) if t[it, k]) + f"; staging-box waits {int(t[it,24])} cyc")
