#!/bin/bash
TAG=${1:-r2n}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log | cut -c1-300
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity_errors.json 2>/dev/null
DN4GL_MLP2_TC=1 timeout 600 python -m pytest tests/test_models_gpu.py tests/test_pipeline_gpu.py -m gpu -q --tb=line -k "counting" > gpurun_out/${TAG}_pytest_mlp2tc.log 2>&1
echo "pytest mlp2tc rc=$?"; tail -8 gpurun_out/${TAG}_pytest_mlp2tc.log | cut -c1-300
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity_errors_mlp2tc.json 2>/dev/null
