set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/dbg_tc2.py > gpurun_out/dbg_tc2.log 2>&1; echo "dbg rc=$?"
cat gpurun_out/dbg_tc2.log | tail -40
