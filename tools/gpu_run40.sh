cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench40.json 2> gpurun_out/bench40.err; echo "bench rc=$?"; tail -5 gpurun_out/bench40.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench40.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
print(json.dumps(d['roofline'])[:900])
print(d.get('roofline_c5')); print(d.get('mlp_stages')); print(d['cpu_baseline'])
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench40_ref.json 2>gpurun_out/bench40_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench40_ref.json
