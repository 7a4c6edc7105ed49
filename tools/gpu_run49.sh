cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/k1_timeline.py > gpurun_out/k1_timeline_c2.txt 2>&1; echo rc=$?; tail -9 gpurun_out/k1_timeline_c2.txt | cut -c1-900
timeout 300 python tools/k1_timeline.py --c5 16384 --dim 64 > gpurun_out/k1_timeline_c5.txt 2>&1; echo rc=$?; tail -5 gpurun_out/k1_timeline_c5.txt | cut -c1-900
