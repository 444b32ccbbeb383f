#!/bin/bash
# DRAM traffic / duration / tensor-pipe activity of the three recon_tc_kernel passes of one dsi_rec call on cfg3 (96x96x60x515).
# Usage: bash tools/gpu/ncu_dsi.sh <tag>
TAG=${1:-dsi}
mkdir -p gpurun_out
cat > /tmp/dsi_once.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch, bench
import fibers_jl_b200 as F
from fibers_jl_b200 import device as D, phantom
dev = torch.device("cuda", 0)
b, g = phantom.dsi_grid_table(); nvox = 96 * 96 * 60; N = b.shape[0]
pitch = (nvox + 63) // 64 * 64
dwi = bench.synth_dwi_device(torch, nvox, b, g, 3, dev, pitch=pitch)
mask = torch.ones(nvox, dtype=torch.uint8, device=dev)
odf = torch.empty((321, pitch), dtype=torch.float32, device=dev); pdf = torch.empty((N, pitch), dtype=torch.float32, device=dev)
peak = [torch.empty((3, pitch), dtype=torch.float32, device=dev) for _ in range(3)]
qa = [torch.empty(pitch, dtype=torch.float32, device=dev) for _ in range(3)]
stats = torch.zeros(2, dtype=torch.int32, device=dev)
plan = D.Plan("dsi", 0, b, g)
for _ in range(3):
    plan.recon(dwi.data_ptr(), pitch, mask.data_ptr(), nvox, pitch, odf.data_ptr(), [p.data_ptr() for p in peak], [q.data_ptr() for q in qa],
               stats.data_ptr(), d_pdf=pdf.data_ptr(), finalize=True)
torch.cuda.synchronize()
PY
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
  --clock-control none -k regex:recon_tc -s 6 -c 3 --csv --log-file gpurun_out/ncu_$TAG.csv python /tmp/dsi_once.py > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log; grep -v "^==" gpurun_out/ncu_$TAG.csv | cut -d, -f5,13- | tail -16
