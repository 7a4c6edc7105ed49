cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_transforms_gpu.py -m gpu -q --tb=short -k "match_weights or build_csr" > gpurun_out/pytest39.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest39.log
tail -25 gpurun_out/pytest39.log | cut -c1-300
