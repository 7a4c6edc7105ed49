cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lin_bwd_kernel" -s 4 -c 2 -o gpurun_out/prof_bwd16 python tools/bench_tc.py --rows 156759 --dims 32 > gpurun_out/ncu_full16.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_full16.log
