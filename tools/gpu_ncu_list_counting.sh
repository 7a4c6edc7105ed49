cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for c in c3 c4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 2500 --csv --log-file gpurun_out/r3t_${c}_launches.csv python tools/bench_counting.py --config $c --steps 6 --warmup 8 > gpurun_out/r3t_${c}_ncu.log 2>&1
  echo "$c rc=$?"
  python tools/launch_summary.py gpurun_out/r3t_${c}_launches.csv 40 > gpurun_out/r3t_${c}_launch_summary.txt 2>&1
  head -32 gpurun_out/r3t_${c}_launch_summary.txt | cut -c1-170; tail -1 gpurun_out/r3t_${c}_launch_summary.txt
done
