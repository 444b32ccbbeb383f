# Same-box A/B of kernel variants: tools/_bin/variants/lib_<name>.so (other builds of the library), each timed twice, interleaved.
for rep in 1 2; do
for v in "$@"; do
  bash tools/gpu/quick_bench.sh FIBERS_CUDA_LIB=$PWD/tools/_bin/variants/lib_$v.so
done; done
