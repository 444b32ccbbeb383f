#!/bin/bash
# Local helper: another build of libfibers_cuda.so with extra flags for recon_tc.cu (same-box A/B through FIBERS_CUDA_LIB).
# Usage: tools/build_variant.sh <name> [nvcc flags ...]   ->  tools/_bin/variants/lib_<name>.so
set -e
name=$1; shift
cd "$(dirname "$0")/.."
python fibers.jl_b200/build.py > /dev/null
mkdir -p tools/_bin/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -pthread "$@" -c fibers.jl_b200/csrc/recon_tc.cu -o /tmp/recon_tc_$name.o
objs=$(ls fibers.jl_b200/build/*.o | grep -v recon_tc.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/_bin/variants/lib_$name.so $objs /tmp/recon_tc_$name.o -lcudart -lcuda -ldl -lz -Xcompiler -pthread
echo tools/_bin/variants/lib_$name.so
