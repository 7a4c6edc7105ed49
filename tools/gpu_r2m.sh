#!/bin/bash
TAG=${1:-r2m}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log | cut -c1-300
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity_errors.json 2>/dev/null
DN4GL_MLP2_TC=1 timeout 600 python -m pytest tests/test_models_gpu.py tests/test_pipeline_gpu.py -m gpu -q --tb=line -k "counting" > gpurun_out/${TAG}_pytest_mlp2tc.log 2>&1
echo "pytest mlp2tc rc=$?"; tail -8 gpurun_out/${TAG}_pytest_mlp2tc.log | cut -c1-300
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity_errors_mlp2tc.json 2>/dev/null
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$? in $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --no-cpu-baseline --no-extras --no-size-hints > gpurun_out/${TAG}_bench_nohints.json 2> gpurun_out/${TAG}_bench_nohints.err
echo "bench nohints rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_nohints.json
# ncu --set full of the three hot kernels (2 launches each), from a short bench run
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lin_fwd_pipe|lin_bwd_pipe|spmm_pipe' -s 30 -c 6 \
    -o gpurun_out/${TAG}_hot python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_hot.log 2>&1
echo "ncu hot rc=$?"
