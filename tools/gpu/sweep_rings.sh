# sweeps the shared-memory budget of recon_tc_kernel<true>: B ring depth, DWI ring depth, ODF staging boxes, L2 prefetch distance
run() { echo "B=$1 D=$2 OBUF=$3 L2PF=$4: $(FIBERS_TC_VERBOSE=1 FIBERS_TC_BSTAGES=$1 FIBERS_TC_DSTAGES=$2 FIBERS_TC_OBUF=$3 FIBERS_TC_L2PF=$4 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/tmp/err.txt | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["roofline"]["kernel_ms"],4), round(d["ms_per_step"],4))') $(grep -m1 'fibers tc' /tmp/err.txt | cut -c12-)"; }
run 4 4 2 0
run 4 4 2 1
run 6 4 1 0
run 6 4 1 1
run 6 4 1 2
run 8 4 0 1
run 6 6 0 1
run 6 4 0 1
run 7 4 0 1
run 5 4 2 1
run 6 2 2 1
run 8 2 1 1
