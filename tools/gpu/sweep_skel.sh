run() { echo "abl=$5 B=$1 D=$2 OBUF=$3 L2PF=$4: $(FIBERS_TC_ABLATE=$5 FIBERS_TC_VERBOSE=1 FIBERS_TC_BSTAGES=$1 FIBERS_TC_DSTAGES=$2 FIBERS_TC_OBUF=$3 FIBERS_TC_L2PF=$4 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/tmp/err.txt | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["roofline"]["kernel_ms"],4), d["clocks"].get("kernel_effective_sm_mhz"))') $(grep -m1 'fibers tc' /tmp/err.txt | cut -c20-)"; }
run 4 4 0 0 15
run 4 6 0 0 15
run 4 8 0 0 15
run 4 4 0 1 15
run 4 4 0 2 15
run 4 8 0 1 15
run 4 4 0 0 7
run 4 8 0 0 7
run 4 8 0 0 6
run 4 8 0 0 2
run 4 8 0 0 0
run 2 8 0 0 0
