#!/bin/bash
# ncu launch list (gpu__time_duration per kernel) of a short bench run -> gpurun_out/<tag>_launches.csv + summary
TAG=${1:-rX}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 900 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 6 --warmup 10 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launch list rc=$?"
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launch_summary.txt 2>&1
head -45 gpurun_out/${TAG}_launch_summary.txt
