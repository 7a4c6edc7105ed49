cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest31.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest31.log
tail -25 gpurun_out/pytest31.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/bench31.json 2> gpurun_out/bench31.err; echo "bench rc=$?"; tail -5 gpurun_out/bench31.err
cut -c1-2500 gpurun_out/bench31.json
