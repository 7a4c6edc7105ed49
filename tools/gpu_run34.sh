cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_models_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest34.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest34.log
tail -12 gpurun_out/pytest34.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/bench34.json 2> gpurun_out/bench34.err; echo "bench rc=$?"; tail -5 gpurun_out/bench34.err
cut -c1-1300 gpurun_out/bench34.json
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench34b.json 2> gpurun_out/bench34b.err; echo "bench rc=$?"; tail -5 gpurun_out/bench34b.err
cut -c1-400 gpurun_out/bench34b.json
