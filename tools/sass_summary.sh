#!/bin/bash
# Instruction histogram per kernel of the shipped library (cuobjdump -sass): which tensor / async-copy / barrier
# instructions each hot kernel really contains.  usage: tools/sass_summary.sh > profiles/r2_sass_summary.txt
LIB=${1:-cugp_b200/libcugp.so}
echo "# cuobjdump -sass $LIB ($(date -u +%F)), nvcc $(nvcc --version | grep release | sed 's/.*release //')"
echo "# columns: DMMA (mma.sync.m8n8k4.f64 = the FP64 tensor instruction of sm_100a; tcgen05 has no f64 kind), LDGSTS (cp.async),"
echo "# UBLKCP (1-D TMA bulk copy), SYNCS (mbarrier), BAR (named / CTA barriers), UCGABAR (cluster barrier), SHFL, MUFU.RSQ64H,"
echo "# DFMA+DMUL+DADD, RED/ATOM (global atomics), LDG/STG, total instructions"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
/Function :/ { if (name != "") emit(); name=$3; delete c; tot=0; next }
/^[ \t]+\/\*[0-9a-f]+\*\// {
  ins=$2; if (ins ~ /^@/) ins=$3;
  tot++;
  if (ins ~ /^DMMA/) c["DMMA"]++;
  else if (ins ~ /^LDGSTS/) c["LDGSTS"]++;
  else if (ins ~ /^UBLKCP/) c["UBLKCP"]++;
  else if (ins ~ /^SYNCS/) c["SYNCS"]++;
  else if (ins ~ /^BAR/) c["BAR"]++;
  else if (ins ~ /^UCGABAR/) c["UCGABAR"]++;
  else if (ins ~ /^SHFL/) c["SHFL"]++;
  else if (ins ~ /^MUFU.RSQ64H/) c["RSQ64H"]++;
  else if (ins ~ /^(DFMA|DMUL|DADD)/) c["FP64"]++;
  else if (ins ~ /^(RED|ATOM)/) c["ATOM"]++;
  else if (ins ~ /^(LDG|STG)/) c["LDGSTG"]++;
}
function emit() {
  short=name; gsub(/_ZN4cugp[0-9]+_GLOBAL__N__[0-9a-f]+_[0-9]+_[a-z_]+_cu_[0-9a-f]+/, "", short);
  printf "%-110s DMMA %5d LDGSTS %4d UBLKCP %2d SYNCS %3d BAR %3d UCGABAR %2d SHFL %4d RSQ64H %3d FP64 %5d ATOM %2d LDG/STG %4d total %6d\n", substr(short,1,110), c["DMMA"], c["LDGSTS"], c["UBLKCP"], c["SYNCS"], c["BAR"], c["UCGABAR"], c["SHFL"], c["RSQ64H"], c["FP64"], c["ATOM"], c["LDGSTG"], tot
}
END { if (name != "") emit() }'
echo "# totals over the library:"
cuobjdump -sass "$LIB" 2>/dev/null | grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?[A-Z0-9_.]+" | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | awk '$2 ~ /^(DMMA|LDGSTS|UBLKCP|SYNCS|UCGABAR|UTCHMMA|UTCQMMA|UTMALDG|HMMA|HGMMA|SHFL|MUFU|DFMA|RED|ATOMG|BAR)$/ {printf "%s %d\n", $2, $1}'
