#!/bin/bash
# One gpurun call that produces everything profiles/ needs for a round (run from the repo root on the GPU box):
#   gpurun --timeout 1500 -- 'bash tools/gpu_suite.sh r2a'
# 1. pytest -m gpu  2. bench.py (N=1)  3. ncu launch list of the bench  4. ncu --set full of the aggregation kernel.
# Copy the files you want judged from gpurun_out/ into profiles/ afterwards (tools/launch_summary.py, tools/ncu_summary.py).
TAG=${1:-rX}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 400 python bench.py > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/${TAG}_bench_c2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 900 --csv \
    --log-file gpurun_out/${TAG}_launches_bench_c2.csv python bench.py --steps 6 --warmup 10 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_pipe -s 40 -c 2 \
    -o gpurun_out/${TAG}_spmm_pipe_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
