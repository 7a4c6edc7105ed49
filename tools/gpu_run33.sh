cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest33.log
tail -6 gpurun_out/pytest33.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/bench33.json 2> gpurun_out/bench33.err; echo "bench rc=$?"; tail -5 gpurun_out/bench33.err
cut -c1-400 gpurun_out/bench33.json
timeout 300 python tools/bench_k1_c2.py > gpurun_out/k1_c2_sweep.jsonl 2>&1; cat gpurun_out/k1_c2_sweep.jsonl | cut -c1-250
