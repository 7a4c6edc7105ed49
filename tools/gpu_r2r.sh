#!/bin/bash
TAG=${1:-r2r}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log | cut -c1-300
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity_errors.json 2>/dev/null
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$? in $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
S=$(date +%s)
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
echo "bench ref rc=$? in $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/${TAG}_bench_ref.json; tail -3 gpurun_out/${TAG}_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
