# frame pitch of a device-resident slab: 16- / 8- / 4-byte aligned rows through 1 / 2 / 4 TMA maps, against the cp.async staging
python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "aligned" 2>&1 | tail -5
BENCH_KERNELS_ONLY=pitch python tools/gpu/bench_kernels.py 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(f\"{d['kernel']:100s} {d['ms']:9.3f} ms {d['frac_of_measured_hbm']*100:5.1f}%\")
" | tee gpurun_out/pitch_ab_r2.txt
