cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in c3 c4; do
  timeout 200 python tools/bench_counting.py --config $c > gpurun_out/bc27_$c.log 2>&1; echo "rc=$?"
  tail -12 gpurun_out/bc27_$c.log | cut -c1-300
done
