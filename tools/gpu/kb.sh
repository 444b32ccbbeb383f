python tools/gpu/bench_kernels.py 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(f\"{d['kernel']:70s} {d['ms']:9.3f} ms {d['voxels_per_s']:.3e} vox/s {d['algorithmic_GBps']:8.1f} GB/s {d['frac_of_measured_hbm']*100:5.1f}%\")
"
