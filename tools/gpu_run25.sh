cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/dbg_steps.py > gpurun_out/dbg_steps25.log 2>&1; echo "rc=$?"
cut -c1-330 gpurun_out/dbg_steps25.log | tail -14
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench25.json 2> gpurun_out/bench25.err; echo "bench rc=$?"
head -c 300 gpurun_out/bench25.json; echo
