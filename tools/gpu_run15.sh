set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_pipeline_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest15.log
tail -30 gpurun_out/pytest15.log | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench15.json 2> gpurun_out/bench15.err; echo "bench rc=$?"
tail -5 gpurun_out/bench15.err
head -c 1500 gpurun_out/bench15.json
