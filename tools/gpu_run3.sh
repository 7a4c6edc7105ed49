set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_agg_gpu.py -m gpu -x -q > gpurun_out/pytest_agg.log 2>&1; echo "pytest agg rc=$?" >> gpurun_out/pytest_agg.log
tail -15 gpurun_out/pytest_agg.log
timeout 600 python tools/agg_sweep.py --graphs 4096,16384,65536 --dims 32,64,128 --modes rows,tiled --smem 200 --warps 16,24,32 --out gpurun_out/sweep3.json > gpurun_out/sweep3.log 2>&1; echo "sweep rc=$?"
timeout 600 python tools/agg_sweep.py --graphs 16384,65536 --dims 256,512 --modes rows,tiled --smem 200 --warps 16 > gpurun_out/sweep3b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_pipe -s 3 -c 2 -o gpurun_out/prof_pipe2_d64 python tools/agg_sweep.py --graphs 16384 --dims 64 --modes tiled --smem 200 --warps 32 --iters 2 > gpurun_out/ncu_full3.log 2>&1
cat gpurun_out/sweep3.log gpurun_out/sweep3b.log
