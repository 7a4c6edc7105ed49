set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu10.log
tail -60 gpurun_out/pytest_gpu10.log | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "bench rc=$?"
tail -5 gpurun_out/bench10.err
head -c 3500 gpurun_out/bench10.json
