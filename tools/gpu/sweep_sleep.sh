for cs in 100 400 1000 0; do for ps in 200 1000 0; do
echo "conv_sleep=$cs prod_sleep=$ps: $(FIBERS_TC_CONV_SLEEP=$cs FIBERS_TC_PROD_SLEEP=$ps python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["roofline"]["kernel_ms"], d["ms_per_step"])')"
done; done
