#!/bin/bash
# Turns the .ncu-rep files a GPU run left in gpurun_out/ into the committed summaries under profiles/ (round tag $1).
set -e
cd "$(dirname "$0")/.."
R=${1:-r2}
extra() { ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c "
import csv,sys,re
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]; r=rows[2]
for a,b,c in zip(h,u,r):
    if re.search(r'sm__pipe_(fma|fmaheavy|alu|xu)_cycles_active.avg.pct_of_peak_sustained_(active|elapsed)|gpu__dram_throughput.avg.pct|sm__issue_active.avg.pct', a): print(f'{a:88s} {b:16s} {c}')
"; }
for pair in "final:cluster" "sk:cluster_sk" "range:range" "decode:decode"; do
  src=gpurun_out/prof_${R}_${pair%%:*}.ncu-rep; name=${pair##*:}
  [ -f "$src" ] || continue
  python tools/ncu_summary.py "$src" > profiles/${R}_${name}_ncu_summary.txt
  extra "$src" >> profiles/${R}_${name}_ncu_summary.txt
  python tools/ncu_lines.py "$src" > profiles/${R}_${name}_lines.txt
done
[ -f gpurun_out/${R}_launches.csv ] && cp gpurun_out/${R}_launches.csv profiles/${R}_launches_bench.csv
echo refreshed
