set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python tools/agg_sweep.py --graphs 1024,4096,16384,65536 --dims 64,128,256,512 --modes rows,tiled --smem 64,96,190 --out gpurun_out/sweep.json > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_tiled -s 6 -c 3 -o gpurun_out/prof_tiled_sweep python tools/agg_sweep.py --graphs 16384 --dims 64 --modes tiled --smem 96 --iters 2 > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
cat gpurun_out/bench.json | head -c 3000
tail -30 gpurun_out/sweep.log
