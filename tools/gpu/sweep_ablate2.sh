for a in 32 48 32 48; do
echo "ablate=$a: $(FIBERS_TC_ABLATE=$a python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("kernel_ms", round(d["roofline"]["kernel_ms"],4), "eff_clock", d["clocks"].get("kernel_effective_sm_mhz"))')"
done
