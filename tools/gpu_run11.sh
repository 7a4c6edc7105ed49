set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/bench_tc.py --out gpurun_out/bench_tc11.json > gpurun_out/bench_tc11.log 2>&1; echo "rc=$?"
cat gpurun_out/bench_tc11.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 700 --csv --log-file gpurun_out/launches11.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu11.log 2>&1; echo "ncu rc=$?"
