set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "bench rc=$?"
tail -3 gpurun_out/bench13.err
head -c 1200 gpurun_out/bench13.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 700 --csv --log-file gpurun_out/launches13.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu13.log 2>&1; echo "ncu rc=$?"
