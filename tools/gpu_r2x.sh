#!/bin/bash
# one GPU call: the -m gpu suite (no -x: every failure is listed), the C2 bench line, the launch list of a few steps
TAG=${1:-r2x}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log | cut -c1-300
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity_errors.json 2>/dev/null
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
bash tools/gpu_ncu_list.sh ${TAG}
