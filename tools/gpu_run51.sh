cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_agg_gpu.py tests/test_models_gpu.py tests/test_pipeline_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest51.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest51.log
tail -8 gpurun_out/pytest51.log | cut -c1-300
timeout 300 python tools/k1_timeline.py --reps 3 > gpurun_out/k1_timeline_c2_c.txt 2>&1; echo rc=$?; tail -5 gpurun_out/k1_timeline_c2_c.txt | cut -c1-900
timeout 300 python tools/bench_k1_c2.py --smem 200 --warps 16,24,32 --dims 32,64 > gpurun_out/k1_c2_sweep_c.jsonl 2>&1; cat gpurun_out/k1_c2_sweep_c.jsonl | cut -c1-250
