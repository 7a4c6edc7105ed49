cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/host_profile_transform.py > gpurun_out/host_prof_transform.txt 2>&1; echo rc=$?
head -5 gpurun_out/host_prof_transform.txt
