#!/bin/bash
# Runs on the GPU box (via gpurun): smoke, both bench arms, ncu launch list + one full capture.
# Usage: bash tools/gpu/run_bench_profile.sh <tag> [kernel]
TAG=${1:-r1}; KERNEL=${2:-auto}
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --kernel $KERNEL > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; tail -c 1500 gpurun_out/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'recon|qa_scale|stats_init|fit_|sample_max|tile_' --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --kernel $KERNEL > gpurun_out/ncu_list_$TAG.log 2>&1
tail -3 gpurun_out/ncu_list_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"recon_$( [ $KERNEL = simt ] && echo simt || echo tc )_kernel" -s 2 -c 1 -o gpurun_out/prof_$TAG \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --kernel $KERNEL > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out
