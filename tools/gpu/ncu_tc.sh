#!/bin/bash
# ncu --set full capture of one recon_tc_kernel launch.  Usage: bash tools/gpu/ncu_tc.sh <tag> [shape]
TAG=${1:-tc}; SHAPE=${2:-145,174,145}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:recon_tc -s 2 -c 1 -o gpurun_out/prof_$TAG \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --kernel tc --shape $SHAPE > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-300
