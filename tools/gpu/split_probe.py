"""One small launch of the tensor-core kernel per DWI alignment case (run under compute-sanitizer when a case faults)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
import fibers_jl_b200 as F
from fibers_jl_b200 import device as D
dev = torch.device("cuda", 0)
bval, bvec = bench.make_tables()
nvox = 50003
pal = (nvox + 63) // 64 * 64
ref = bench.synth_dwi_device(torch, nvox, bval, bvec, 5, dev, pitch=pal)
mask = torch.ones(nvox, dtype=torch.uint8, device=dev)
D.set_kernel("tc")
plan = D.Plan("gqi", 0, bval, bvec, F.sphere_642, 1.25)
first = None
for dp in [pal] + [int(x) for x in sys.argv[1:]]:
    dwi = torch.zeros((bval.shape[0], dp), dtype=torch.float32, device=dev)
    dwi[:, :nvox] = ref[:, :nvox]
    odf = torch.zeros((321, pal), dtype=torch.float32, device=dev)
    peak = [torch.zeros((3, pal), dtype=torch.float32, device=dev) for _ in range(3)]
    qa = [torch.zeros(pal, dtype=torch.float32, device=dev) for _ in range(3)]
    stats = torch.zeros(2, dtype=torch.int32, device=dev)
    plan.recon(dwi.data_ptr(), dp, mask.data_ptr(), nvox, pal, odf.data_ptr(), [p.data_ptr() for p in peak], [q.data_ptr() for q in qa], stats.data_ptr(), finalize=True)
    torch.cuda.synchronize()
    if first is None: first = odf.clone()
    print("pitch", dp, "ok; equal to the aligned result:", bool(torch.equal(first, odf)), flush=True)
