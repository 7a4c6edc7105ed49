cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest19.log
tail -30 gpurun_out/pytest19.log | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench19.json 2> gpurun_out/bench19.err; echo "bench rc=$?"
tail -3 gpurun_out/bench19.err
head -c 300 gpurun_out/bench19.json; echo
