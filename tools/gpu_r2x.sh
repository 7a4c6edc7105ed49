#!/bin/bash
TAG=${1:-r2x}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log | cut -c1-300
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity_errors.json 2>/dev/null
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
