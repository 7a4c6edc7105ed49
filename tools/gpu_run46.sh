cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest46.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest46.log
tail -8 gpurun_out/pytest46.log | cut -c1-300
for c in c3 c4; do
  timeout 200 python tools/bench_counting.py --config $c --steps 50 --warmup 20 > gpurun_out/bc46_$c.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bc46_$c.log | cut -c1-400
  timeout 200 python tools/bench_counting.py --config $c --steps 50 --warmup 20 --opt torch --no-overlap > gpurun_out/bc46_${c}_old.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bc46_${c}_old.log | cut -c1-400
done
