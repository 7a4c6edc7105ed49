cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest29.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest29.log
tail -8 gpurun_out/pytest29.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/bench29.json 2> gpurun_out/bench29.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/bench29.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke29.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke29.log
