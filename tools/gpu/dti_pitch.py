"""dti_fit / adc_fit (device-resident, cfg4 shape) against the frame pitch of the DWI slab and of the outputs."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, bench
import fibers_jl_b200 as F
from fibers_jl_b200 import device as D
dev = torch.device("cuda", 0)
shape = (145, 174, 145); nvox = int(np.prod(shape))
bval, bvec = bench.make_tables(); N = 288
HBM = bench.measured_peaks()[0]
def timeit(fn, steps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps
mask = torch.ones(nvox, dtype=torch.uint8, device=dev)
plan = D.Plan("dti", 0, bval, bvec)
big = os.environ.get("DTI_PITCH_BALLAST")           # allocate 6 GB first, like bench.py's ODF / peak buffers
ballast = torch.empty(int(6e9) // 4, dtype=torch.float32, device=dev) if big else None
for dp, op in ((nvox, nvox), (nvox + 2, nvox + 2), ((nvox + 63) // 64 * 64, nvox), (nvox, (nvox + 63) // 64 * 64), ((nvox + 63) // 64 * 64,) * 2, (nvox + 31, nvox + 31)):
    dwi = bench.synth_dwi_device(torch, nvox, bval, bvec, int(os.environ.get("DTI_PITCH_SEED", "3")), dev, pitch=dp)
    outs = [torch.empty((n, op), dtype=torch.float32, device=dev) for n in (1, 1, 1, 1, 3, 3, 3, 1, 1, 1)]
    ms = timeit(lambda: plan.dti_fit(dwi.data_ptr(), dp, mask.data_ptr(), nvox, op, [o.data_ptr() for o in outs]))
    print(f"dwi pitch {dp} (row {dp * 4 % 128:3d} mod 128 B), out pitch {op}: {ms:.4f} ms  {nvox * (4 * N + 65) / ms / 1e6 / HBM:.4f} of the HBM roof", flush=True)
    del dwi, outs
