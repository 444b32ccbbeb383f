# A/B on ONE box: round-1 tree (_r1/, a git worktree of the round-1 commit) against the current tree, alternating runs
f() { python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("kernel_ms", round(d["roofline"]["kernel_ms"],4), "step_ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), "eff_clock", d["clocks"].get("kernel_effective_sm_mhz"), d["clocks"].get("reasons"))'; }
for steps in 10 20 100; do
  for rep in 1 2; do
    echo "r1  steps=$steps: $(cd _r1 && python bench.py --steps $steps --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | f)"
    echo "new steps=$steps: $(python bench.py --steps $steps --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | f)"
  done
done
