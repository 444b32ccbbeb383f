#!/bin/bash
# Round-end regression on one B200: GPU suite, smoke, both bench arms, kernel table, ncu launch list of one bench step.
TAG=${1:-r2final}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 1500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; tail -c 600 gpurun_out/bench_ref_$TAG.json
python tools/gpu/bench_kernels.py 2>/dev/null | grep "^{" > gpurun_out/kernels_$TAG.jsonl; wc -l gpurun_out/kernels_$TAG.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'recon|qa_scale|stats_init|fit_|sample_max|tile_' --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-dti > gpurun_out/ncu_list_$TAG.log 2>&1
tail -2 gpurun_out/ncu_list_$TAG.log | cut -c1-200
