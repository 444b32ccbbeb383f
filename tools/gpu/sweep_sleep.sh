# back-off (ns) of the converters' A-slot wait and of the two producers' ring waits: the polling loops are 15 % of the executed
# instructions of recon_tc_kernel (ncu source page), so the poll period trades issue slots against wake-up latency
for cfg in "100 200" "300 200" "1000 200" "100 1000" "300 1000" "1000 1000" "2000 2000" "0 0"; do set -- $cfg
bash tools/gpu/quick_bench.sh FIBERS_TC_CONV_SLEEP=$1 FIBERS_TC_PROD_SLEEP=$2
done
