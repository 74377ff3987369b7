#!/bin/bash
# usage: tools/ncu_digest.sh gpurun_out/prof_X   (reads prof_X.ncu-rep, writes prof_X_raw.csv / prof_X_source.csv, prints a digest)
p=$1
ncu -i $p.ncu-rep --page raw --csv > ${p}_raw.csv 2>/dev/null
ncu -i $p.ncu-rep --page source --csv --print-source cuda,sass > ${p}_source.csv 2>/dev/null
python tools/ncu_summary.py ${p}_raw.csv | grep -E "gpu__time_duration.sum|registers_per_thread  |sm__warps_active.avg.pct|smsp__inst_executed.sum |thread_inst_executed_per_inst|issue_active.avg.pct|pipe_(fma|alu|xu|fp64)\.avg\.pct_of_peak_sustained_active|dram__bytes_(read|write).sum " 
python tools/ncu_summary.py ${p}_raw.csv | grep -E "issue_stalled.*per_issue_active" | awk '{print $1, $3}' | sed 's/smsp__average_warps_issue_stalled_//; s/_per_issue_active.ratio//' | sort -k2 -n -r | head -8
