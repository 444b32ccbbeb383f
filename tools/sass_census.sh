#!/bin/bash
# Mnemonic census + tensor / TMA / TMEM instruction sites of the shipped recon_tc_kernel<kTma = 1, kTrace = false> (one DWI map:
# 16-byte aligned rows) and the TMA sites of the 2- / 4-map instantiations (8- / 4-byte aligned rows).
# Runs in the build container (cuobjdump only):  bash tools/sass_census.sh > profiles/r2_recon_tc.sass
SO=${1:-fibers.jl_b200/libfibers_cuda.so}
cuobjdump -sass $SO > /tmp/_all.sass
start=$(grep -n "Function : .*recon_tc_kernelILi1ELb0" /tmp/_all.sass | cut -d: -f1)
end=$(awk -v s=$start 'NR>s && /Function :/ {print NR; exit}' /tmp/_all.sass)
sed -n "${start},${end}p" /tmp/_all.sass | grep -v "^\s*/\* 0x" | sed 's/\/\* 0x[0-9a-f]* \*\///' > /tmp/_k.sass
echo "# recon_tc_kernel<kTma = 1, kTrace = false> (sm_100a) as shipped in $SO -- mnemonic census and the tensor / TMA / TMEM instruction sites"
echo "# produced by: bash tools/sass_census.sh (cuobjdump -sass)"
echo; echo "## census"
for m in UTCHMMA UTMALDG UTMASTG UTMAPF LDTM STTM UTCBAR UTCATOMSWS "SYNCS.PHASECHK" "SYNCS.ARRIVE" LDGSTS VIMNMX3 FMNMX3 PRMT "LDS" "STS" "STG" "LDG" FFMA FMUL F2FP MUFU "BAR.SYNC" NANOSLEEP ELECT R2UR ATOMS SHFL; do
  printf "%-16s %s\n" "$m" "$(grep -c "[ .]$m" /tmp/_k.sass)"
done
echo "total instructions $(grep -c '/\*[0-9a-f]*\*/' /tmp/_k.sass)"
echo; echo "## sites"
grep -n "UTCHMMA\|UTMALDG\|UTMASTG\|UTMAPF\|LDTM\|STTM\|UTCBAR\|UTCATOMSWS" /tmp/_k.sass | cut -c1-140
for k in 2 4; do
  start=$(grep -n "Function : .*recon_tc_kernelILi${k}ELb0" /tmp/_all.sass | cut -d: -f1)
  end=$(awk -v s=$start 'NR>s && /Function :/ {print NR; exit}' /tmp/_all.sass)
  echo; echo "## recon_tc_kernel<kTma = $k>: DWI slab through $k tensor maps (rows not 16-byte aligned) -- TMA load sites"
  sed -n "${start},${end}p" /tmp/_all.sass | grep -v "^\s*/\* 0x" | sed 's/\/\* 0x[0-9a-f]* \*\///' | grep -n "UTMALDG\|UTMAPF" | cut -c1-140
done
