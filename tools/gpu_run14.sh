cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/host_profile.py > gpurun_out/host_profile14.log 2>&1; echo "rc=$?"
cut -c1-220 gpurun_out/host_profile14.log | head -90
