set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_pipe -s 40 -c 1 -o gpurun_out/prof_pipe_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
tail -3 gpurun_out/ncu_c2.log | cut -c1-300
