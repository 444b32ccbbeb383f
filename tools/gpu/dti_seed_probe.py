import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
dev = torch.device("cuda", 0)
nvox = 145*174*145
bval, bvec = bench.make_tables()
for seed in (1000, 3):
    d = bench.synth_dwi_device(torch, nvox, bval, bvec, seed, dev)
    neg = (d <= 0).sum(0)
    print("seed", seed, "max dropped per voxel", int(neg.max()), "hist", torch.bincount(neg.to(torch.int64))[:10].tolist(),
          "zeros", int((d == 0).sum()), "nonfinite", int((~torch.isfinite(d)).sum()), "min positive", float(d[d > 0].min()), "max", float(d.max()))
    b0 = torch.tensor(bval == bval.min(), device=dev)
    print("   voxels with all min-b samples dropped:", int(((d[b0] <= 0).sum(0) == int(b0.sum())).sum()))
