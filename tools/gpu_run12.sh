set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_tc_gpu.py tests/test_models_gpu.py -m gpu -q --tb=short > gpurun_out/pytest_tc12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc12.log
tail -15 gpurun_out/pytest_tc12.log | cut -c1-250
timeout 600 python tools/bench_tc.py --rows 156759 --dims 32 --out gpurun_out/bench_tc12.json > gpurun_out/bench_tc12.log 2>&1; echo "rc=$?"
cat gpurun_out/bench_tc12.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lin_fwd_kernel|lin_bwd_kernel" -s 8 -c 4 -o gpurun_out/prof_tc12 python tools/bench_tc.py --rows 156759 --dims 32 > gpurun_out/ncu_full12.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_full12.log
