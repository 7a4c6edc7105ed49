#!/bin/bash
TAG=${1:-r2t}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_counting_loss_gpu.py tests/test_mlp_tc_gpu.py -m gpu -q --tb=line > gpurun_out/${TAG}_pytest_a.log 2>&1
echo "pytest default rc=$?"; tail -4 gpurun_out/${TAG}_pytest_a.log | cut -c1-300
DN4GL_MLP2_TC=0 timeout 600 python -m pytest tests/test_zz_counting_loss_gpu.py -m gpu -q --tb=line > gpurun_out/${TAG}_pytest_b.log 2>&1
echo "pytest mlp2tc=0 rc=$?"; tail -4 gpurun_out/${TAG}_pytest_b.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest all rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log | cut -c1-300
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity_errors.json 2>/dev/null
