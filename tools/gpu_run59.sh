cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
STEPS=4000 timeout 300 python tools/e2e_stall.py > gpurun_out/e2e_stall.txt 2>&1; echo rc=$?; head -40 gpurun_out/e2e_stall.txt | cut -c1-250
