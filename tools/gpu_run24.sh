cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_tc_gpu.py tests/test_models_gpu.py tests/test_pipeline_gpu.py -m gpu -q --tb=short > gpurun_out/pytest24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest24.log
tail -6 gpurun_out/pytest24.log | cut -c1-250
timeout 300 python tools/dbg_tc.py 2>&1 | tail -4
timeout 300 python tools/bench_tc.py --rows 156759,1000000 --dims 32,64 --out gpurun_out/bench_tc24.json > gpurun_out/bench_tc24.log 2>&1
grep -E "bwd" gpurun_out/bench_tc24.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench24.json 2> gpurun_out/bench24.err; echo "bench rc=$?"
head -c 300 gpurun_out/bench24.json; echo
