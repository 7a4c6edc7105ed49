#!/bin/bash
TAG=${1:-r2j}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CS=$PWD/dummynode4graphlearning_b200/csrc
timeout 600 python -m pytest tests/test_mlp_tc_gpu.py -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest_mlp.log 2>&1
echo "pytest mlp rc=$?"; tail -12 gpurun_out/${TAG}_pytest_mlp.log | cut -c1-250
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest_models.log 2>&1
echo "pytest models rc=$?"; tail -5 gpurun_out/${TAG}_pytest_models.log | cut -c1-250
timeout 300 python tools/bench_tc.py --dims 32 > gpurun_out/${TAG}_bench_tc_pipe.jsonl 2> gpurun_out/${TAG}_bench_tc_pipe.err
echo "bench_tc pipe rc=$?"; grep -v addmm gpurun_out/${TAG}_bench_tc_pipe.jsonl
DN4GL_LIB=$CS/libdn4gl_pipetl.so timeout 300 python tools/pipe_timeline.py --dims 32 > gpurun_out/${TAG}_pipe_tl.jsonl 2> gpurun_out/${TAG}_pipe_tl.err
echo "pipe tl rc=$?"; tail -3 gpurun_out/${TAG}_pipe_tl.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
