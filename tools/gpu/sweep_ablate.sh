# timing ablations of recon_tc_kernel (results are wrong by design when a bit is set): 1 = one MMA product instead of three,
# 2 = no local-maximum scan, 4 = no ODF stores / keys in the drain, 8 = no conversion math / LDS in the converters
for a in 0 1 2 4 8 3 6 12 7 15; do
echo "ablate=$a: $(FIBERS_TC_ABLATE=$a python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("kernel_ms", round(d["roofline"]["kernel_ms"],4), "eff_clock", d["clocks"].get("kernel_effective_sm_mhz"))')"
done
