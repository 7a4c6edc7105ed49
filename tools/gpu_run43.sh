cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest43.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest43.log
tail -6 gpurun_out/pytest43.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke43.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke43.log
timeout 400 python bench.py > gpurun_out/bench43.json 2> gpurun_out/bench43.err; echo "bench rc=$?"; tail -3 gpurun_out/bench43.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 900 --csv --log-file gpurun_out/r1f_launches.csv python bench.py --steps 6 --warmup 10 --no-cpu-baseline > gpurun_out/ncu_l43.log 2>&1
echo "launchlist rc=$?"; tail -2 gpurun_out/ncu_l43.log | cut -c1-200
