cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/dbg_tc.py > gpurun_out/dbg_tc22.log 2>&1; tail -18 gpurun_out/dbg_tc22.log
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest22.log
tail -12 gpurun_out/pytest22.log | cut -c1-250
timeout 300 python tools/bench_tc.py --rows 156759,1000000 --dims 32,64 > gpurun_out/bench_tc22.log 2>&1
grep -E "fwd_plain|fwd_bn|bwd_bn" gpurun_out/bench_tc22.log
