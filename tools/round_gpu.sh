#!/bin/bash
# usage (on the GPU box): tools/round_gpu.sh <tag> [quick] -- the round's measurement pass in one gpurun call:
# bench (both arms) -> gpurun_out/<tag>_bench_n1*.json, the -m gpu suite -> <tag>_gpu_tests.log, and the ncu launch list of
# the env-only loop -> <tag>_launches_env_steady.csv. Every leg runs under its own timeout so one hang cannot eat the call.
tag=${1:-rXX}
quick=${2:-}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 420 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${tag}_bench_n1.err
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -5 gpurun_out/${tag}_gpu_tests.log
if [ -n "$quick" ]; then  # intermediate pass: env-only timing, launch list, tests; no reference arm / smoke / net profile
  timeout 240 bash tools/launch_list.sh ${tag}_launches_env_steady > gpurun_out/${tag}_launch_cycle.txt 2>&1
  tail -32 gpurun_out/${tag}_launch_cycle.txt
  python tools/profile_env.py --cycles 40 2>&1 | tail -2
  head -c 1500 gpurun_out/${tag}_bench_n1.json
  exit 0
fi
timeout 300 python bench.py --impl reference > gpurun_out/${tag}_bench_n1_reference.json 2> gpurun_out/${tag}_bench_ref.err
echo "reference arm rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 240 bash tools/launch_list.sh ${tag}_launches_env_steady > gpurun_out/${tag}_launch_cycle.txt 2>&1
echo "launch list rc=$?"; tail -40 gpurun_out/${tag}_launch_cycle.txt
timeout 120 python tools/profile_net.py > gpurun_out/${tag}_profile_net.txt 2>&1
echo "profile_net rc=$?"; tail -15 gpurun_out/${tag}_profile_net.txt
head -c 1500 gpurun_out/${tag}_bench_n1.json
