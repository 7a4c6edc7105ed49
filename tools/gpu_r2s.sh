#!/bin/bash
TAG=${1:-r2s}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log | cut -c1-300
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity_errors.json 2>/dev/null
