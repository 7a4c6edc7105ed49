set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mlp_tc_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest_tc12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc12.log
tail -15 gpurun_out/pytest_tc12.log | cut -c1-250
timeout 300 python tools/bench_tc.py --rows 156759,1000000 --dims 32,64 --out gpurun_out/bench_tc12.json > gpurun_out/bench_tc12.log 2>&1; echo "rc=$?"
cat gpurun_out/bench_tc12.log
