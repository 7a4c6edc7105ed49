cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_transforms_gpu.py -m gpu -q --tb=short -k "empty_inputs or without_gradient" > gpurun_out/pytest48.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest48.log
tail -30 gpurun_out/pytest48.log | cut -c1-300
