cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest53.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest53.log
tail -12 gpurun_out/pytest53.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke53.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke53.log
for i in 1 2 3; do
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench53_$i.json 2> gpurun_out/bench53_$i.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench53_$i.json'))
print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['breakdown']['cudaMalloc_calls_in_timed_region'])"
done
