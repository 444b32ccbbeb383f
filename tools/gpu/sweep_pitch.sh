# frame pitch of the device-resident slabs (DWI and outputs) against DRAM channel mapping: rows must stay 16-byte aligned for TMA
for cfg in "64 0" "4 0" "64 4" "64 8" "64 16" "64 32" "64 36" "64 68" "1024 0"; do set -- $cfg
echo -n "align $1 extra $2: "; BENCH_PITCH_ALIGN=$1 BENCH_PITCH_EXTRA=$2 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("gqi", round(d["roofline"]["kernel_ms"],4), round(d["roofline"]["frac"],4), "dti", round(d["dti_fit"]["ms_per_step"],4), round(d["dti_fit"]["roofline"]["frac"],4), "adc", round(d["adc_fit"]["ms_per_step"],4), round(d["adc_fit"]["roofline"]["frac"],4))'
done
