cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest18.log
tail -8 gpurun_out/pytest18.log | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench18.json 2> gpurun_out/bench18.err; echo "bench rc=$?"
head -c 300 gpurun_out/bench18.json; echo
timeout 300 python tools/host_profile.py > gpurun_out/host_profile18.log 2>&1
head -8 gpurun_out/host_profile18.log
