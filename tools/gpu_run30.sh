cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# launch list: skip the 10 warm-up steps' worth of launches roughly, then take 900 launches (~3 steps incl. graph replays)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 900 --csv --log-file gpurun_out/r1d_launches.csv python bench.py --steps 6 --warmup 10 --no-cpu-baseline > gpurun_out/ncu_l30.log 2>&1
echo "launchlist rc=$?"; tail -2 gpurun_out/ncu_l30.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_pipe -s 40 -c 2 -o gpurun_out/r1d_spmm_pipe_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f30.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_f30.log | cut -c1-300
ls -la gpurun_out | tail -5
