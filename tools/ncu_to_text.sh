#!/bin/bash
# ncu_to_text.sh <report.ncu-rep> <object file under carmpc_b200/_lib/obj> <mangled kernel substring> <out.txt>
# Raw-metric summary + per-source-line profile of one capture as text (the .ncu-rep itself is too large to keep).
rep=$1; obj=$2; pat=$3; out=$4
tmp=$(mktemp -d)
( cd $tmp && cuobjdump -xelf all $obj > /dev/null 2>&1 )
cubin=$(ls $tmp/*.cubin | head -1)
{ echo "== ncu --set full --clock-control none: raw metrics"; python tools/ncu_summary.py $rep;
  echo; echo "== per source line (share of executed instructions / of stall samples)"; python tools/ncu_lines.py $rep $cubin $pat 45; } > $out 2>&1
rm -rf $tmp
