cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/step_split.py > gpurun_out/step_split.txt 2>&1; echo rc=$?; cat gpurun_out/step_split.txt | tail -8
