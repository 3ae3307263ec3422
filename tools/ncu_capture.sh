#!/bin/bash
# usage: tools/ncu_capture.sh <kernel-regex> <skip> <name> -- runs one `ncu --set full` capture of the env-only loop
# (tools/profile_env.py) and leaves two small text files in gpurun_out/: <name>.raw.csv (all metrics of the launch)
# and <name>.lines.txt (per-source-line instruction / stall shares). The .ncu-rep itself is deleted (tens of MB).
set -e
k=$1; sk=$2; nm=$3
rep=/tmp/$nm.ncu-rep
ncu --set full --clock-control none --import-source on -k "regex:$k" -s $sk -c 1 -o /tmp/$nm python tools/profile_env.py --cycles 4 > /dev/null 2>&1
ncu -i $rep --page raw --csv > gpurun_out/$nm.raw.csv 2>/dev/null
ncu -i $rep --page source --csv --print-source cuda,sass 2>/dev/null | python tools/ncu_lines.py 25 > gpurun_out/$nm.lines.txt
rm -f $rep
