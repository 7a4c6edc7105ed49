cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench37.json 2> gpurun_out/bench37.err; echo "bench rc=$?"; tail -5 gpurun_out/bench37.err
cut -c1-300 gpurun_out/bench37.json
python -c "
import json; d=json.load(open('gpurun_out/bench37.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'])"
