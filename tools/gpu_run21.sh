cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest21.log
tail -30 gpurun_out/pytest21.log | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench21.json 2> gpurun_out/bench21.err; echo "bench rc=$?"
head -c 300 gpurun_out/bench21.json; echo
