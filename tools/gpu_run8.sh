set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu8.log
tail -8 gpurun_out/pytest_gpu8.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench rc=$?"
head -c 3000 gpurun_out/bench8.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 800 --csv --log-file gpurun_out/launches8.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu8.log 2>&1; echo "ncu rc=$?"
