#!/bin/bash
# usage: tools/launch_list.sh <name> -- ncu launch list (gpu__time_duration.sum per launch) of the env-only loop, written to
# gpurun_out/<name>.csv; prints the per-kernel times of the last full steady-state cycle (tools/launch_cycle.py).
nm=${1:-launches}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/$nm.csv python tools/profile_env.py --cycles 4 > /dev/null 2>&1
python tools/launch_cycle.py gpurun_out/$nm.csv
