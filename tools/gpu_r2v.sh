#!/bin/bash
TAG=${1:-r2v}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$? in $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
