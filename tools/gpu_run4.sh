set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_agg_gpu.py -m gpu -x -q > gpurun_out/pytest_agg.log 2>&1; echo "pytest agg rc=$?" >> gpurun_out/pytest_agg.log
tail -15 gpurun_out/pytest_agg.log
timeout 600 python tools/agg_sweep.py --graphs 4096,16384,65536 --dims 32,64,128 --modes tiled --smem 200 --warps 16,24,32 --out gpurun_out/sweep4.json > gpurun_out/sweep4.log 2>&1; echo "sweep rc=$?"
timeout 600 python tools/agg_sweep.py --graphs 16384,65536 --dims 256,512 --modes tiled --smem 200 --warps 16 > gpurun_out/sweep4b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_pipe -s 3 -c 1 -o gpurun_out/prof_pipe3_d64 python tools/agg_sweep.py --graphs 16384 --dims 64 --modes tiled --smem 200 --warps 32 --iters 2 > gpurun_out/ncu_full4.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench rc=$?"
cat gpurun_out/sweep4.log gpurun_out/sweep4b.log
cat gpurun_out/bench4.json | head -c 3500
