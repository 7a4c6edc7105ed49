#!/bin/bash
TAG=${1:-r2l}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log | cut -c1-300
DN4GL_MLP2_TC=1 timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -q --tb=line -k "counting" > gpurun_out/${TAG}_pytest_mlp2tc.log 2>&1
echo "pytest mlp2tc rc=$?"; tail -12 gpurun_out/${TAG}_pytest_mlp2tc.log | cut -c1-300
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$? in $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
