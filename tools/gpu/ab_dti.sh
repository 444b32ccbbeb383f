# Same-box A/B of DTI / ADC kernel variants (tools/_bin/variants/lib_<name>.so)
for rep in 1 2; do for v in "$@"; do
  echo "== $v"; BENCH_KERNELS_ONLY=dti FIBERS_CUDA_LIB=$PWD/tools/_bin/variants/lib_$v.so python tools/gpu/bench_kernels.py 2>/dev/null | python -c 'import sys,json
for l in sys.stdin:
    d=json.loads(l); print("  ", d["kernel"], round(d["ms"],4), "ms", round(d["frac_of_measured_hbm"],4))'
done; done
