#!/bin/bash
# ncu launch list (gpu__time_duration per kernel) of a few bench steps -> gpurun_out/<tag>_launches.csv + summary.
# Per-launch times are cold-cache and serialised: compare SHARES and COUNTS, not absolutes.
TAG=${1:-rX}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3400 -c 420 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 30 --no-cpu-baseline --no-extras \
    > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launch list rc=$?"
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv 70 > gpurun_out/${TAG}_launch_summary.txt 2>&1
head -75 gpurun_out/${TAG}_launch_summary.txt
