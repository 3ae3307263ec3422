#!/bin/bash
# usage (on the GPU box): tools/round_ncu.sh <tag> -- the round's ncu evidence in one gpurun call, all small text files in gpurun_out/:
#   <tag>_traffic.csv            dram bytes + time per launch of the env-only loop  -> tools/ncu_traffic.py -> profiles/ncu_traffic.json
#   <tag>_k_*.raw.csv/.lines.txt `ncu --set full` of the top kernels                -> tools/ncu_summary.py -> profiles/<tag>_ncu_*.md
#   <tag>_net_launches.csv       launch list of the policy-net forward              -> tools/launch_hist.py
# With two DP lanes the host launches lane 0's kernels, then lane 1's: per cycle the matching launches of a kernel come lane by lane.
tag=${1:-rXX}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_traffic.csv python tools/profile_env.py --cycles 5 > /dev/null 2>&1
echo "traffic rc=$?"
# ncu matches the function name without template arguments. k_sp_eval: 8 launches per lane and cycle (levels 7..0: W0, D0, W1, ...),
# 16 per cycle; the W1 level of lane 0 in the 3rd cycle = skip 2 * 16 + 2
timeout 200 bash tools/ncu_capture.sh k_sp_eval 34 ${tag}_k_sp_eval_W1; echo "eval rc=$?"
# k_sp_expand: 8 per lane and cycle (levels 0..7: D3, W3, D2, W2, D1, W1, D0, W0); W1 of lane 0 in the 3rd cycle = skip 2 * 16 + 5
timeout 200 bash tools/ncu_capture.sh k_sp_expand 37 ${tag}_k_sp_expand_W1; echo "expand rc=$?"
timeout 200 bash tools/ncu_capture.sh "k_encode_store" 2 ${tag}_k_encode_store; echo "store rc=$?"
timeout 200 bash tools/ncu_capture.sh "k_sp_finalize" 4 ${tag}_k_sp_finalize; echo "finalize rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_net_launches.csv \
    python tools/profile_net.py > /dev/null 2>&1
echo "net rc=$?"
python tools/launch_hist.py gpurun_out/${tag}_net_launches.csv 0 | tail -25
ls -la gpurun_out | tail -20
