cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest23.log
tail -12 gpurun_out/pytest23.log | cut -c1-250
timeout 300 python tools/dbg_tc.py 2>&1 | tail -6
timeout 300 python tools/bench_tc.py --rows 156759,1000000 --dims 32,64 --out gpurun_out/bench_tc23.json > gpurun_out/bench_tc23.log 2>&1
grep -E "fwd_plain|fwd_bn|bwd_bn" gpurun_out/bench_tc23.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench23.json 2> gpurun_out/bench23.err; echo "bench rc=$?"
head -c 300 gpurun_out/bench23.json; echo
