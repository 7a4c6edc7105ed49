cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/dbg_steps.py > gpurun_out/dbg_steps20.log 2>&1; echo "rc=$?"
cut -c1-400 gpurun_out/dbg_steps20.log | tail -20
