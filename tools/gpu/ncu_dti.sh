#!/bin/bash
# ncu --set full capture of the DTI fit kernel on the HCP-shaped volume.  Usage: bash tools/gpu/ncu_dti.sh <tag>
TAG=${1:-dti}
mkdir -p gpurun_out
cat > /tmp/dti_once.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch, bench
import fibers_jl_b200 as F
from fibers_jl_b200 import device as D
dev = torch.device("cuda", 0)
b, g = bench.make_tables(); nvox = 145 * 174 * 145
dwi = bench.synth_dwi_device(torch, nvox, b, g, 3, dev)
mask = torch.ones(nvox, dtype=torch.uint8, device=dev)
outs = [torch.empty((n, nvox), dtype=torch.float32, device=dev) for n in (1, 1, 1, 1, 3, 3, 3, 1, 1, 1)]
plan = D.Plan("dti", 0, b, g)
for _ in range(3):
    plan.dti_fit(dwi.data_ptr(), nvox, mask.data_ptr(), nvox, nvox, [o.data_ptr() for o in outs])
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fit_full -s 1 -c 1 -o gpurun_out/prof_$TAG python /tmp/dti_once.py > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
