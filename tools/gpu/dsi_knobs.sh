# DSI (cfg3) against the depth of the raw DWI ring and the L2 prefetch of the DWI boxes
one() { env "$@" BENCH_KERNELS_ONLY=dsi python tools/gpu/bench_kernels.py 2>/dev/null | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(' '.join(sys.argv[1:]) or 'default', round(d['ms'],3), 'ms')" "$@"; }
for rep in 1 2; do
one A=1; one FIBERS_TC_DSTAGES=4; one FIBERS_TC_DSTAGES=6; one FIBERS_TC_L2PF=1
done 2>&1 | tee gpurun_out/dsi_knobs_r2.txt
