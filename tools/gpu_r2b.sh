#!/bin/bash
# round 2, call b: the pipelined forward stage (mlp_pipe.cu) -- parity, micro-benchmark old vs new, bench line;
# + the one-gY-copy descriptor experiment on the (old) backward kernel
TAG=${1:-r2b}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CS=$PWD/dummynode4graphlearning_b200/csrc
timeout 600 python -m pytest tests/test_mlp_tc_gpu.py -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest_mlp.log 2>&1
echo "pytest mlp rc=$?"; tail -15 gpurun_out/${TAG}_pytest_mlp.log | cut -c1-250
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest_models.log 2>&1
echo "pytest models rc=$?"; tail -5 gpurun_out/${TAG}_pytest_models.log | cut -c1-250
DN4GL_LIB=$CS/libdn4gl_exp_onegcopy.so DN4GL_LIN_SERIAL=1 timeout 600 python -m pytest tests/test_mlp_tc_gpu.py -m gpu -q --tb=line -k "bwd or chain or mlp2" > gpurun_out/${TAG}_pytest_onegcopy.log 2>&1
echo "pytest onegcopy rc=$?"; tail -6 gpurun_out/${TAG}_pytest_onegcopy.log | cut -c1-250
timeout 300 python tools/bench_tc.py > gpurun_out/${TAG}_bench_tc_pipe.jsonl 2> gpurun_out/${TAG}_bench_tc_pipe.err
echo "bench_tc pipe rc=$?"; grep fwd gpurun_out/${TAG}_bench_tc_pipe.jsonl
DN4GL_LIN_SERIAL=1 timeout 300 python tools/bench_tc.py > gpurun_out/${TAG}_bench_tc_serial.jsonl 2> gpurun_out/${TAG}_bench_tc_serial.err
echo "bench_tc serial rc=$?"; grep fwd gpurun_out/${TAG}_bench_tc_serial.jsonl
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
