cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -q --tb=short > gpurun_out/pytest36.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest36.log
tail -25 gpurun_out/pytest36.log | cut -c1-300
timeout 300 python tools/host_profile_transform.py > gpurun_out/host_prof_step.txt 2>&1; echo rc=$?
grep -n "host" gpurun_out/host_prof_step.txt | head
