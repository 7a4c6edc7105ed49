cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mlp_tc_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest17.log
tail -8 gpurun_out/pytest17.log | cut -c1-250
timeout 300 python tools/bench_tc.py --rows 156759,1000000 --dims 32,64 --out gpurun_out/bench_tc17.json > gpurun_out/bench_tc17.log 2>&1; echo "rc=$?"
grep -v "bn_act\|torch_addmm" gpurun_out/bench_tc17.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "bench rc=$?"
head -c 300 gpurun_out/bench17.json
