cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q --tb=short > gpurun_out/r4a_gemm_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r4a_gemm_pytest.log | cut -c1-250
