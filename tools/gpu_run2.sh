set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_agg_gpu.py -m gpu -x -q > gpurun_out/pytest_agg.log 2>&1; echo "pytest agg rc=$?" >> gpurun_out/pytest_agg.log
tail -15 gpurun_out/pytest_agg.log
timeout 600 python tools/agg_sweep.py --graphs 1024,4096,16384,65536 --dims 64,128,256,512 --modes rows,tiled --smem 200 --out gpurun_out/sweep2.json > gpurun_out/sweep2.log 2>&1; echo "sweep rc=$?"
timeout 300 python tools/agg_sweep.py --graphs 16384 --dims 64,512 --modes tiled --smem 200,150,100 --known 0 > gpurun_out/sweep2_unknown.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_agg_gpu.py > gpurun_out/pytest_rest.log 2>&1; echo "pytest rest rc=$?" >> gpurun_out/pytest_rest.log
tail -5 gpurun_out/pytest_rest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_pipe -s 3 -c 2 -o gpurun_out/prof_pipe_d64 python tools/agg_sweep.py --graphs 16384 --dims 64 --modes tiled --smem 200 --iters 2 > gpurun_out/ncu_full2.log 2>&1
cat gpurun_out/sweep2.log gpurun_out/sweep2_unknown.log
cat gpurun_out/bench2.json | head -c 3500
