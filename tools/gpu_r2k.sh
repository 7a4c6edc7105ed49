#!/bin/bash
TAG=${1:-r2k}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CS=$PWD/dummynode4graphlearning_b200/csrc
for NE in 1 2; do
export DN4GL_BWD_NEPI=$NE
timeout 600 python -m pytest tests/test_mlp_tc_gpu.py -m gpu -x -q --tb=short -k "bwd or chain or mlp2" > gpurun_out/${TAG}_pytest_mlp_$NE.log 2>&1
echo "nepi=$NE pytest mlp rc=$?"; tail -3 gpurun_out/${TAG}_pytest_mlp_$NE.log | cut -c1-250
DN4GL_LIB=$CS/libdn4gl_pipetl.so timeout 300 python tools/pipe_timeline.py --dims 32 > gpurun_out/${TAG}_pipe_tl_$NE.jsonl 2> gpurun_out/${TAG}_pipe_tl_$NE.err
echo "pipe tl rc=$?"; tail -3 gpurun_out/${TAG}_pipe_tl_$NE.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_$NE.json 2> gpurun_out/${TAG}_bench_$NE.err
echo "bench rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_$NE.json; tail -3 gpurun_out/${TAG}_bench_$NE.err
done
