cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench52.json 2> gpurun_out/bench52.err; echo "bench rc=$?"; tail -3 gpurun_out/bench52.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench52.json'))
print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']))
r=d['roofline']; print(r['avg_launch_us'], r['frac'], r['cold_l2_single_launch_us'], r['in_step_eager_us'])
print(d['roofline_c5'])
PY
timeout 300 python tools/agg_sweep.py --graphs 4096,16384,65536 --dims 64,128,256 --modes tiled > gpurun_out/agg_sweep52.jsonl 2>&1; cut -c1-220 gpurun_out/agg_sweep52.jsonl | tail -12
