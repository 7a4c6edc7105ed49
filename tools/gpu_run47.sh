cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/dbg_flat_vs_torch.py > gpurun_out/dbg47.log 2>&1; echo rc=$?; tail -8 gpurun_out/dbg47.log | cut -c1-400
