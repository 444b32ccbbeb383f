#!/bin/bash
# dynamic instruction count + duration of one recon_tc_kernel launch (1 ncu pass).  Usage: bash tools/gpu/ncu_inst.sh
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:recon_tc -s 2 -c 1 \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --kernel tc 2>&1 | grep -E "inst_executed|time_duration|issue_active"
