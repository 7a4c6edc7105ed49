#!/bin/bash
TAG=${1:-r2q}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CS=$PWD/dummynode4graphlearning_b200/csrc
timeout 900 python -m pytest tests/test_zzz_conj_direct_gpu.py tests/test_zzz_hints_gpu.py -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
DN4GL_LIB=$CS/libdn4gl_pdl1.so timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_pdl1.json 2> gpurun_out/${TAG}_bench_pdl1.err
echo "bench pdl1 rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_pdl1.json; tail -5 gpurun_out/${TAG}_bench_pdl1.err
DN4GL_LIB=$CS/libdn4gl_pdl1.so timeout 900 python -m pytest tests/test_mlp_tc_gpu.py tests/test_pipeline_gpu.py tests/test_agg_gpu.py -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest_pdl1.log 2>&1
echo "pytest pdl1 rc=$?"; tail -5 gpurun_out/${TAG}_pytest_pdl1.log | cut -c1-300
