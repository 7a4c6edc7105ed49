#!/bin/bash
# A/B of programmatic dependent launch (csrc/common.cuh: DN_LAUNCH / DN_PDL_WAIT) against the product build, one gpurun call:
#   make -C dummynode4graphlearning_b200/csrc libdn4gl_pdl1.so libdn4gl_pdl2.so     (here, before the call: the .so travel)
#   gpurun --timeout 1500 -- 'bash tools/gpu_ab_pdl.sh r2a'
# For each variant: 1. parity -- every GPU test that goes through the converted kernels (aggregation, MLP stages, models,
# pipelines incl. the CUDA-graph train step) with DN4GL_LIB pointing at the variant; 2. the C2 bench line (no CPU leg).
# A variant is only worth keeping if its parity run is green AND ms_per_step drops; outputs under gpurun_out/<tag>_pdl*.
TAG=${1:-rX}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CS=$PWD/dummynode4graphlearning_b200/csrc
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_pdl_bench_base.json 2> gpurun_out/${TAG}_pdl_bench_base.err
echo "base bench rc=$?"; cut -c1-260 gpurun_out/${TAG}_pdl_bench_base.json
for V in 1 2; do
  LIB=$CS/libdn4gl_pdl$V.so
  test -f "$LIB" || { echo "build libdn4gl_pdl$V.so first"; continue; }
  DN4GL_LIB=$LIB timeout 900 python -m pytest tests/test_agg_gpu.py tests/test_mlp_tc_gpu.py tests/test_models_gpu.py \
      tests/test_pipeline_gpu.py -m gpu -x -q --tb=short > gpurun_out/${TAG}_pdl${V}_pytest.log 2>&1
  echo "pdl$V parity rc=$?"; tail -2 gpurun_out/${TAG}_pdl${V}_pytest.log | cut -c1-200
  DN4GL_LIB=$LIB timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_pdl${V}_bench.json 2> gpurun_out/${TAG}_pdl${V}_bench.err
  echo "pdl$V bench rc=$?"; cut -c1-260 gpurun_out/${TAG}_pdl${V}_bench.json
done
